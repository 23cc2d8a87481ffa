"""GPU parity at layer and model level against the reference's golden fixtures (tests/golden)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, atol, rtol=1e-4, what=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    tol = atol + rtol * ref.abs().max().item()
    assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("mode", ["fused_chunk", "chunk", "fused_recurrent"])
@pytest.mark.parametrize("sc", [False, True])
def test_layer_cfg1(golden_layer, mode, sc):
    """BASELINE config 1: GatedLinearAttention d256 h4 T128 B1 vs the CPU torch reference (fla off)."""
    from lina_speech_b200.model import GatedLinearAttention
    from lina_speech_b200.fla_api import Cache
    g, p = golden_layer, ("sc_" if sc else "nosc_")
    layer = GatedLinearAttention(mode=mode, hidden_size=256, num_heads=4, use_short_conv=sc, layer_idx=0).eval()
    layer.load_state_dict({k[len(p) + 2:]: v for k, v in g.items() if k.startswith(p + "w.")})
    layer = layer.to(DEV)
    x = g[p + "x"].to(DEV)
    with torch.no_grad():
        _close(layer(x), g[p + "y"], 2e-5, what="one shot")
        cache = Cache()
        cache.update(layer.init_state(1), 0, offset=0)
        _close(layer(x[:, :100], past_key_values=cache, use_cache=True), g[p + "y_pre"], 2e-5, what="prefill")
    with torch.inference_mode():
        ys = torch.cat([layer(x[:, t:t + 1], past_key_values=cache, use_cache=True) for t in range(100, 128)], 1)
    _close(ys, g[p + "y_steps"], 2e-5, what="steps")
    for i, s in enumerate(cache.states[0]):
        _close(s, g[p + f"state{i}"], 1e-4, what=f"state{i}")


def _tiny(golden_model):
    import lina_speech_b200.model as m
    rnn = m.AttentiveGLA(64, 2, 2, blind=True, use_short_conv=True, pos_type="convolutional")
    lm = m.LinaModel(rnn, 64, 1, 64, 3, 3, 32, txt_encoder=m.TextEncoder(64, 2, n_layers=1, dropout=0.0, rotary=False))
    lm.load_state_dict({k[2:]: v for k, v in golden_model.items() if k.startswith("w.")})
    return lm.to(DEV).eval()


def test_tiny_model_forward(golden_model):
    g, lm = golden_model, _tiny(golden_model)
    with torch.no_grad():
        logits, loss, att, _, _ = lm(g["x"].to(DEV), g["y"].to(DEV), g["enc_mask"].to(DEV), g["ca_mask"].to(DEV),
                                     logits_mask=g["y_mask"].to(DEV))
    _close(logits, g["logits"], 2e-4, what="logits")
    _close(loss, g["loss"], 1e-4, what="loss")
    _close(att, g["att"], 1e-4, what="att")


def test_tiny_model_forward_with_initial_state_and_its_gradient(golden_model):
    """initial-state tuning entry point (initial_state.py:114-128): forward in train mode + fused_recurrent
    with a rank-1 state; loss matches the reference and the state factors receive gradients (dh0 path)."""
    g, lm = golden_model, _tiny(golden_model)
    lm.attentive_rnn.to_mode("fused_recurrent")
    lm.train()
    params = [(torch.nn.Parameter(g[f"tune_k{i}"].to(DEV)), torch.nn.Parameter(g[f"tune_v{i}"].to(DEV))) for i in range(4)]
    st = lm.attentive_rnn.get_state_from_params(params, 2, scale=0.02)
    logits, loss, _, _, _ = lm(g["x"].to(DEV), g["y"].to(DEV), g["enc_mask"].to(DEV), g["ca_mask"].to(DEV),
                               logits_mask=g["y_mask"].to(DEV), init_state=st)
    _close(logits, g["logits_init_state"], 2e-4, what="logits(init_state)")
    _close(loss, g["loss_init_state"], 1e-4, what="loss(init_state)")
    loss.backward()
    for kp, vp in params:
        assert kp.grad is not None and vp.grad is not None and torch.isfinite(kp.grad).all()
        assert kp.grad.abs().max() > 0


@pytest.mark.parametrize("graph", [False, True])
def test_tiny_model_greedy_generation_bit_exact(golden_model, graph):
    g, lm = golden_model, _tiny(golden_model)
    qs, atts, stop_tokens, cuts = lm.generate_batch(g["xt"].to(DEV), batch_size=3, prompt=g["prompt"].to(DEV),
                                                    max_seqlen=24, k=1, force_max_seqlen=True, cuda_graph=graph)
    assert torch.equal(qs.cpu(), g["qs"]), "greedy token ids differ from the reference"
    _close(atts, g["atts"], 1e-4, what="atts")
    assert len(cuts) == 3


def test_step_loop_equals_forward(golden_model):
    """AttentiveGLA.step x T == AttentiveGLA.forward on the same tokens (cache plumbing incl. pos_net)."""
    lm = _tiny(golden_model)
    rnn = lm.attentive_rnn
    torch.manual_seed(0)
    B, T, Tx = 2, 20, 9
    x, ctx = torch.randn(B, T, 64, device=DEV), torch.randn(B, Tx, 64, device=DEV)
    with torch.inference_mode():
        y, att = rnn(x, ctx)
        cache = rnn.init_state(batch_size=B)
        ys, atts = zip(*[rnn.step(x[:, t:t + 1], ctx, t, cache)[:2] for t in range(T)])
    _close(torch.cat(ys, 1), y, 1e-4, what="step loop y")
    _close(torch.cat(atts, 2), att, 1e-4, what="step loop att")


def test_initial_state_tuning_loop_reduces_the_loss(golden_model):
    """train_initial_state (initial_state.py:85-160): Adam on the rank-1 state factors only, through the dh0 path."""
    from lina_speech_b200.initial_state import train_initial_state, speaker_state_dict
    lm = _tiny(golden_model)

    class Tok:
        def encode(self, s):
            return [1] + [3 + (ord(c) % 29) for c in s[5:-5]] + [2]

    torch.manual_seed(0)
    data = [{"audio_token": torch.randint(0, 64, (1, 30)), "text": "hello world"} for _ in range(2)]
    w0 = {k: v.clone() for k, v in lm.state_dict().items()}
    params, losses = train_initial_state(lm, data, Tok(), n_samples=48, lr=0.05, grad_acc=2, batch_size=2, rank=1)
    assert len(losses) == 24 and all(l == l for l in losses)
    assert sum(losses[-4:]) / 4 < sum(losses[:4]) / 4, (losses[:4], losses[-4:])       # it learns the two utterances
    assert all(torch.equal(w0[k], v) for k, v in lm.state_dict().items())                # weights untouched
    sd = speaker_state_dict(params)
    assert set(sd) == {f"layer{i}_{s}" for i in range(4) for s in "kv"} and not lm.training


@pytest.mark.parametrize("graph", [False, True])
def test_prompt_prefill_equals_token_by_token_teacher_forcing(golden_model, graph):
    """generate_batch(prefill_prompt=True): start token + prompt in ONE multi-token pass with the cache (chunkwise GLA, conv
    tails, pos_net state) must continue exactly like the reference's token-by-token teacher forcing (model/modeling_lina.py:
    152-179): same greedy ids (also the golden ones), same attention maps, same stop flags."""
    g, lm = golden_model, _tiny(golden_model)
    kw = dict(batch_size=3, prompt=g["prompt"].to(DEV), max_seqlen=24, k=1, force_max_seqlen=True, cuda_graph=graph)
    qs0, atts0, st0, _ = lm.generate_batch(g["xt"].to(DEV), **kw)
    qs1, atts1, st1, cuts1 = lm.generate_batch(g["xt"].to(DEV), prefill_prompt=True, **kw)
    assert torch.equal(qs1, qs0) and torch.equal(qs1.cpu(), g["qs"]), "prefilled prompt changes the greedy continuation"
    _close(atts1, atts0, 1e-4, what="atts (prefill vs steps)")
    assert torch.equal(st1, st0) and len(cuts1) == 3


def test_train_lina_mirror_steps_reduce_the_loss():
    """train_lina.py:72-120 through the Lightning-free mirror: collate -> step -> backward -> AdamW + cosine warm-up."""
    import lina_speech_b200.model as m
    from lina_speech_b200.train_lina import TrainLina
    from lina_speech_b200.tuning import simple_collate

    class Tok:
        def encode(self, s):
            return [1] + [3 + (ord(c) % 29) for c in s[5:-5]] + [2]

    torch.manual_seed(0)
    tl = TrainLina(m.AttentiveGLA(64, 1, 2, blind=True, use_short_conv=True, pos_type="convolutional"), 64, [0], 64, 3, 3, 40,
                   txt_encoder=m.TextEncoder(64, 2, n_layers=1, dropout=0.0, rotary=False), learning_rate=3e-3,
                   n_warmup_steps=2, n_training_steps=40).to(DEV).train()
    data = [{"audio_token": torch.randint(0, 64, (1, 40)), "text": t} for t in ("hello world", "good morning")]
    batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in simple_collate(data, Tok()).items()}
    batch["crossatt_pos"] = None
    (opt,), (sch,) = tl.configure_optimizers()
    losses = []
    for i in range(20):
        opt.zero_grad(set_to_none=True)
        loss = tl.training_step(batch, i)
        loss.backward()
        opt.step()
        sch["scheduler"].step()
        losses.append(float(loss))
    assert all(l == l for l in losses) and losses[-1] < 0.7 * losses[0], losses
