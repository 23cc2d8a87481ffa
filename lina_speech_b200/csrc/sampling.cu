// Fused top-k sampling of the decode loop (model/tools.py:38-44 topk_sampling + model/modeling_lina.py:159-165):
//   kth = k-th largest logit of the row (UNSCALED, the reference's quirk) ; keep x/temp >= kth ; p = softmax(kept) ;
//   id  = inverse-CDF sample of p with the caller's uniform u in [0,1)
// One CTA per row: the row is staged in shared memory as fp32, the k-th value found by a 32-step bitwise search on the
// order-preserving integer image of the floats (values in registers, one block barrier per step), the softmax sum by a block
// reduction, the sample by a block scan over per-thread segment sums.  Replaces ~15 torch launches per step (topk = sort, div, compare, masked_fill, softmax,
// multinomial ...: ~130 us at bs128) with one.  torch.multinomial's own algorithm cannot be matched bit for bit; the
// distribution is the same, and k = 1 (greedy) returns the arg-max exactly (exact ties are split by u, like multinomial).
#include "common.cuh"

namespace {

constexpr int ST = 256;                 // threads per row
constexpr int MAXV = 8192;              // row length staged in shared memory

__device__ __forceinline__ uint32_t f2ord(float f) {      // monotone float -> uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

template <typename T>
__global__ void __launch_bounds__(ST)
topk_sample_kernel(const T *__restrict__ logits, long long ld, int Vn, int k, float inv_temp, const float *__restrict__ uni,
                   long long *__restrict__ out) {
    __shared__ float row[MAXV];
    __shared__ int wcnt[2][ST / 32];
    __shared__ float red[ST / 32];
    __shared__ int wfirst[ST / 32], wlast[ST / 32];
    __shared__ uint32_t sel_prefix;
    __shared__ float bcast[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T *r = logits + (size_t)blockIdx.x * ld;
    for (int i = tid; i < Vn; i += ST) row[i] = to_f(r[i]);
    __syncthreads();
    // ---- k-th largest: the largest threshold (in the order-preserving integer image) with #{x >= T} >= k, built bit by bit from the
    // top.  Each thread keeps its <= MAXV / ST values in registers; one iteration = that many compares, a warp sum and ONE block
    // barrier (the per-warp counts are double-buffered).  32 iterations ~ 2 us; the 4-pass histogram select this replaces spent
    // most of the kernel's 34 us serialising shared-memory atomics on the two or three bins a row's top byte falls into.
    constexpr int PER = MAXV / ST;
    uint32_t ov[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int i = tid + j * ST;
        ov[j] = i < Vn ? f2ord(row[i]) : 0u;          // 0 sorts below the image of every float: never counted (the threshold is > 0)
    }
    uint32_t thr_o = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = thr_o | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) c += (ov[j] >= cand) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) wcnt[bit & 1][warp] = c;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < ST / 32; ++w) tot += wcnt[bit & 1][w];
        if (tot >= k) thr_o = cand;
    }
    if (tid == 0) sel_prefix = thr_o;
    __syncthreads();
    const float kth = ord2f(sel_prefix);
    // ---- softmax over the kept entries (x * inv_temp >= kth) ----
    float mx = -INFINITY;
    for (int i = tid; i < Vn; i += ST) mx = fmaxf(mx, row[i] * inv_temp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
        for (int w = 1; w < ST / 32; ++w) m = fmaxf(m, red[w]);
        bcast[0] = m;
    }
    __syncthreads();
    mx = bcast[0];                       // max of the scaled row
    // temp != 1 can push every scaled logit below the unscaled k-th value (the reference then feeds NaNs to
    // torch.multinomial, which raises); keep at least the maximum
    const float thr = fminf(kth, mx);
    __syncthreads();
    // thread t owns the contiguous segment [t*per, (t+1)*per): weights written back into `row`
    const int per = (Vn + ST - 1) / ST;
    const int lo = tid * per, hi = min(Vn, lo + per);
    float ssum = 0.f;
    for (int i = lo; i < hi; ++i) {
        const float x = row[i] * inv_temp;
        const float w = (x >= thr) ? __expf(x - mx) : 0.f;
        row[i] = w;
        ssum += w;
    }
    // block-wide scan of the 256 segment sums: the sample falls into the FIRST non-empty segment whose cumulative weight exceeds
    // u * total (the last non-empty one when rounding leaves none), then that segment's owner walks its <= 33 entries
    float incl = ssum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    float wprefix = 0.f, tot = 0.f;
#pragma unroll
    for (int w = 0; w < ST / 32; ++w) { const float t = red[w]; tot += t; if (w < warp) wprefix += t; }
    const float ex = wprefix + incl - ssum;                  // weight of the segments before this thread's
    const float target = uni[blockIdx.x] * tot;
    const unsigned nonempty = __ballot_sync(0xffffffffu, ssum > 0.f);
    const unsigned hit = __ballot_sync(0xffffffffu, ssum > 0.f && ex + ssum > target);
    if (lane == 0) {
        wfirst[warp] = hit ? warp * 32 + __ffs(hit) - 1 : ST;
        wlast[warp] = nonempty ? warp * 32 + 31 - __clz(nonempty) : -1;
    }
    __syncthreads();
    int first = ST, last = -1;
#pragma unroll
    for (int w = 0; w < ST / 32; ++w) { first = min(first, wfirst[w]); last = max(last, wlast[w]); }
    const int win = first < ST ? first : last;
    if (tid == win) {
        float acc = ex;
        int pick = -1;
        for (int i = lo; i < hi; ++i) {
            const float w = row[i];
            if (w > 0.f) { pick = i; if (acc + w > target) break; acc += w; }
        }
        out[blockIdx.x] = (long long)pick;
    }
}

}  // namespace

extern "C" int lina_topk_sample(const void *logits, long long ld, int B, int Vn, int k, float temp, const float *uniform,
                                int64_t *out, int dtype, void *stream) {
    LINA_REQUIRE(logits && uniform && out && B > 0 && Vn > 0 && ld >= Vn && temp > 0.f, LINA_ERR_BAD_ARG,
                 "topk_sample: bad argument");
    LINA_REQUIRE(lina_dtype_ok(dtype), LINA_ERR_BAD_ARG, "topk_sample: unknown dtype");
    LINA_REQUIRE(Vn <= MAXV, LINA_ERR_UNSUPPORTED, "topk_sample: vocabulary %d > %d", Vn, MAXV);
    LINA_REQUIRE(k >= 1 && k <= Vn, LINA_ERR_BAD_ARG, "topk_sample: k=%d outside [1, %d]", k, Vn);
    LINA_DISPATCH_DTYPE(dtype, topk_sample_kernel<T_><<<B, ST, 0, (cudaStream_t)stream>>>(
                                   (const T_ *)logits, ld, Vn, k, 1.f / temp, uniform, (long long *)out));
    LINA_LAUNCH_OK("topk_sample_kernel");
    return LINA_OK;
}
