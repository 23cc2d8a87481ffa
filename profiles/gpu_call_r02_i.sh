#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_skinny_gpu.py -m gpu -q -x > gpurun_out/r02i_skinny.log 2>&1; echo "skinny rc=$?"; tail -4 gpurun_out/r02i_skinny.log
timeout 600 python - <<'PY'
import sys, torch, json
sys.path.insert(0,'.')
import bench
import lina_speech_b200.model.base_blocks as BB
dev=torch.device('cuda',0)
lm=bench.build_model(dev, torch.bfloat16)
x,_,_,_=bench.synth_inputs(1,8,128,0)
xt=x[0].to(dev)
for flag in (True, False):
    BB.SKINNY_STEP=flag
    for B in (8, 32):
        tm={}
        lm.generate_batch(xt,batch_size=B,max_seqlen=8,k=100,force_max_seqlen=True,cuda_graph=True)
        lm.generate_batch(xt,batch_size=B,max_seqlen=400,k=100,force_max_seqlen=True,cuda_graph=True,stop_check_interval=1<<30,_timing=tm)
        torch.cuda.synchronize()
        print('skinny' if flag else 'library', 'B', B, 'ms/step', tm['start'].elapsed_time(tm['end'])/tm['steps'])
PY
