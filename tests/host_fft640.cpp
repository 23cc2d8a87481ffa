// Host harness for lina_speech_b200/csrc/fft640.cuh: runs the per-lane phases of the ISTFT warp kernel one lane after the
// other (the kernel separates them with __syncwarp()), so the index arithmetic of the device code is checked on the CPU.
// Built by tests/test_host.py with g++; test infrastructure only.
#include <vector>
#include "../lina_speech_b200/csrc/fft640.cuh"

extern "C" int fft640_frames(const float *h, const float *window, float *frames, int nframes) {
    using namespace fft640;
    std::vector<float> tab(TABLE_FLOATS), scratch(WARP_FLOATS);
    for (int i = 0; i < TABLE_FLOATS; ++i) tab[i] = table_entry(i, window);
    float *re0 = scratch.data(), *im0 = re0 + NBINS, *re1 = im0 + NBINS, *im1 = re1 + M;
    for (int f = 0; f < nframes; ++f) {
        const float *hf = h + (size_t)f * (N + 2);
        for (int lane = 0; lane < 32; ++lane) phase_polar(lane, hf, re0, im0);
        for (int lane = 0; lane < 32; ++lane) phase_r10(lane, re0, im0, tab.data(), re1, im1);
        for (int lane = 0; lane < 32; ++lane) phase_r4<10>(lane, re1, im1, tab.data(), T_ST2, re0, im0);
        for (int lane = 0; lane < 32; ++lane) phase_r4<40>(lane, re0, im0, tab.data(), T_ST3, re1, im1);
        for (int lane = 0; lane < 32; ++lane) phase_r4_out(lane, re1, im1, tab.data(), frames + (size_t)f * N);
    }
    return 0;
}
