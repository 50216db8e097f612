// (T)FEP estimator and bootstrap kernels.
//
//   tfepb_lse            single-pass online log-sum-exp over the work values  (analysis/estimator.py:61-86)
//   tfepb_exp_table      e_i = exp(v_i - max) once, so the resampling loop is gather + add
//   tfepb_bootstrap_sums per-resample sum of gathered e_i                      (analysis/bootstrap.py:185-233)
//   tfepb_mt19937_*      the index stream torch.randint draws on a CPU generator (bootstrap.py:214-218)
//
// All are HBM / L2 bound: lse reads 4 (8) bytes per sample once; bootstrap_sums reads one index
// (or generates it from Philox4x32-10 in registers) and one table entry per draw.
#include "common.cuh"

namespace tfepb {
namespace {

constexpr int LSE_THREADS = 256;
constexpr int LSE_MAX_BLOCKS = 2048;

struct MS {
    double m, s;
};

__device__ __forceinline__ MS ms_combine(MS a, MS b) {
    if (b.s == 0.0) return a;
    if (a.s == 0.0) return b;
    const double m = a.m > b.m ? a.m : b.m;
    return MS{m, a.s * exp(a.m - m) + b.s * exp(b.m - m)};
}

__device__ __forceinline__ MS ms_block_reduce(MS v) {
    __shared__ double sm[LSE_THREADS / 32], ss[LSE_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MS u{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
        v = ms_combine(v, u);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[warp] = v.m; ss[warp] = v.s; }
    __syncthreads();
    if (warp == 0) {
        v = lane < LSE_THREADS / 32 ? MS{sm[lane], ss[lane]} : MS{0.0, 0.0};
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            MS u{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
            v = ms_combine(v, u);
        }
    }
    return v;   // valid in thread 0
}

// Each thread keeps a running (max, sum) over groups of values: one exp per value plus one rescale per
// group, the sum carried in double.  The bulk is read with 16-byte vector loads, four of them in flight per
// thread (the kernel is HBM bound: 4 or 8 bytes per sample, read once); the unaligned head / tail is scalar.
// fp32 data is reduced in the log2 domain (values pre-multiplied by log2(e), hardware ex2: one FFMA + one MUFU
// + one FADD per sample, so that the instruction stream stays below the memory time); fp64 data uses exp.
template <typename T> struct LseDom;
template <> struct LseDom<float> {
    static __device__ __forceinline__ float factor() { return 1.4426950408889634f; }
    static __device__ __forceinline__ float exp(float d) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
        return r;
    }
    static __device__ __forceinline__ double to_natural(float m) { return (double)m * 0.6931471805599453094; }
};
template <> struct LseDom<double> {
    static __device__ __forceinline__ double factor() { return 1.0; }
    static __device__ __forceinline__ double exp(double d) { return ::exp(d); }
    static __device__ __forceinline__ double to_natural(double m) { return m; }
};

template <typename T, int N>
__device__ __forceinline__ void lse_absorb(const T (&v)[N], int count, T& m, double& s, bool& have) {
    if (count <= 0) return;
    T gm = v[0];
#pragma unroll
    for (int u = 1; u < N; ++u)
        if (u < count) gm = v[u] > gm ? v[u] : gm;
    if (!have || gm > m) {
        if (have) s *= (double)LseDom<T>::exp(m - gm);
        m = gm;
        have = true;
    }
    T part = T(0);
#pragma unroll
    for (int u = 0; u < N; ++u)
        if (u < count) part += LseDom<T>::exp(v[u] - m);
    s += (double)part;
}

template <typename T>
__global__ void __launch_bounds__(LSE_THREADS, 4) lse_partial_kernel(const T* __restrict__ w, const T* __restrict__ logw,
                                                                  int64_t n, T scale_nat, double* __restrict__ partials) {
    constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte load
    constexpr int LOADS = 4;                      // vector loads in flight per thread
    struct alignas(16) Pack { T v[VEC]; };
    const T dom = LseDom<T>::factor();            // v = dom * (scale w + logw): log2 domain for fp32
    const T scale = scale_nat * dom;
    T m = T(0);
    double s = 0.0;
    bool have = false;
    const int64_t nthreads = (int64_t)gridDim.x * LSE_THREADS;
    const int64_t t0 = (int64_t)blockIdx.x * LSE_THREADS + threadIdx.x;
    // elements before the first 16-byte boundary of w (logw, if present, is only vector-loaded when co-aligned)
    int64_t head = ((16 - (reinterpret_cast<uintptr_t>(w) & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    const bool vec_ok = logw == nullptr || ((reinterpret_cast<uintptr_t>(logw + head) & 15) == 0);
    const int64_t nvec = vec_ok ? (n - head) / VEC : 0;
    const Pack* wv = reinterpret_cast<const Pack*>(w + head);
    const Pack* lv = reinterpret_cast<const Pack*>(logw ? logw + head : nullptr);
    // full groups: every load of the group is in range, no per-value bookkeeping
    int64_t base = t0;
    for (; base + (int64_t)(LOADS - 1) * nthreads < nvec; base += nthreads * LOADS) {
        Pack a[LOADS], b[LOADS];
#pragma unroll
        for (int k = 0; k < LOADS; ++k) {
            a[k] = wv[base + (int64_t)k * nthreads];
            if (logw) b[k] = lv[base + (int64_t)k * nthreads];
        }
        T v[LOADS * VEC];
#pragma unroll
        for (int k = 0; k < LOADS; ++k)
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[k * VEC + e] = logw ? scale * a[k].v[e] + dom * b[k].v[e] : scale * a[k].v[e];
        T gm = v[0];
#pragma unroll
        for (int u = 1; u < LOADS * VEC; ++u) gm = fmax(gm, v[u]);
        if (!have || gm > m) {
            if (have) s *= (double)LseDom<T>::exp(m - gm);
            m = gm;
            have = true;
        }
        T p0 = T(0), p1 = T(0);
#pragma unroll
        for (int u = 0; u < LOADS * VEC; u += 2) {
            p0 += LseDom<T>::exp(v[u] - m);
            p1 += LseDom<T>::exp(v[u + 1] - m);
        }
        s += (double)(p0 + p1);
    }
    // ragged last group
    for (; base < nvec; base += nthreads * LOADS) {
        Pack a[LOADS], b[LOADS];
#pragma unroll
        for (int k = 0; k < LOADS; ++k) {
            const int64_t i = base + (int64_t)k * nthreads;
            if (i < nvec) {
                a[k] = wv[i];
                if (logw) b[k] = lv[i];
            }
        }
        T v[LOADS * VEC];
        int count = 0;
#pragma unroll
        for (int k = 0; k < LOADS; ++k) {
            const int64_t i = base + (int64_t)k * nthreads;
            if (i < nvec) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[k * VEC + e] = scale * a[k].v[e] + (logw ? dom * b[k].v[e] : T(0));
                count = (k + 1) * VEC;
            }
        }
        lse_absorb<T, LOADS * VEC>(v, count, m, s, have);
    }
    // scalar head and tail
    const int64_t done = head + nvec * VEC;
    for (int64_t i = t0; i < head + (n - done); i += nthreads) {
        const int64_t j = i < head ? i : done + (i - head);
        T v1[1] = {scale * w[j] + (logw ? dom * logw[j] : T(0))};
        lse_absorb<T, 1>(v1, 1, m, s, have);
    }
    MS r = ms_block_reduce(MS{LseDom<T>::to_natural(m), have ? s : 0.0});
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = r.m;
        partials[2 * blockIdx.x + 1] = r.s;
    }
}

__global__ void __launch_bounds__(LSE_THREADS) lse_final_kernel(const double* __restrict__ partials, int nparts,
                                                                double* __restrict__ out2) {
    MS v{0.0, 0.0};
    for (int i = threadIdx.x; i < nparts; i += LSE_THREADS) v = ms_combine(v, MS{partials[2 * i], partials[2 * i + 1]});
    v = ms_block_reduce(v);
    if (threadIdx.x == 0) {
        out2[0] = v.m;
        out2[1] = v.s;
    }
}

// The same final reduction, and the estimate itself: result = -kT ((max + log sum) - log_norm) in the dtype of the data
// (what analysis/estimator.py:75-86 evaluates with five scalar tensor operations after the logsumexp).
template <typename T>
__global__ void __launch_bounds__(LSE_THREADS) lse_final_estimate_kernel(const double* __restrict__ partials, int nparts,
                                                                         double* __restrict__ out2, double neg_kT, double log_norm,
                                                                         T* __restrict__ result) {
    MS v{0.0, 0.0};
    for (int i = threadIdx.x; i < nparts; i += LSE_THREADS) v = ms_combine(v, MS{partials[2 * i], partials[2 * i + 1]});
    v = ms_block_reduce(v);
    if (threadIdx.x == 0) {
        out2[0] = v.m;
        out2[1] = v.s;
        result[0] = (T)(neg_kT * ((v.m + log(v.s)) - log_norm));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) exp_table_kernel(const T* __restrict__ w, int64_t n, T scale,
                                                        const double* __restrict__ max_dev, float* __restrict__ e) {
    const T m = (T)max_dev[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        e[i] = (float)Math<T>::exp(scale * w[i] - m);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter = draw index / 4, key = seed
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

constexpr int BS_THREADS = 256;
constexpr int BS_DRAWS_PER_THREAD = 32;     // multiple of 4
constexpr int BS_CHUNK = BS_THREADS * BS_DRAWS_PER_THREAD;

// grid = (chunks of BS_CHUNK draws, resamples).  Explicit indices (exact MT19937 stream) or Philox.
// A rank that holds only the shard [shard_lo, shard_lo + shard_len) of the data walks the SAME global index
// stream and adds the draws that fall into its shard; the per-resample sums of all ranks add up.
__device__ __forceinline__ float shard_fetch(const float* __restrict__ e, uint32_t i, uint32_t shard_lo, uint32_t shard_len) {
    const uint32_t j = i - shard_lo;                 // wraps for i < shard_lo -> fails the range test
    return j < shard_len ? __ldg(e + j) : 0.f;
}

__global__ void __launch_bounds__(BS_THREADS) bootstrap_sums_kernel(const float* __restrict__ e, uint32_t max_idx,
                                                                    const int32_t* __restrict__ idx, int64_t ldidx,
                                                                    int64_t sample_size, uint64_t seed, uint64_t offset,
                                                                    uint32_t shard_lo, uint32_t shard_len,
                                                                    const int64_t* __restrict__ sizes,
                                                                    double* __restrict__ out) {
    const int r = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * BS_CHUNK;
    // per-resample number of draws (stratified resampling: multinomial counts per table tile); the Philox counters
    // keep the stride of the largest one
    const int64_t stride4 = (sample_size + 3) / 4;
    if (sizes != nullptr) {
        sample_size = sizes[r];
        if (j0 >= sample_size) return;
    }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (idx != nullptr) {
        const int32_t* row = idx + (int64_t)r * ldidx;
#pragma unroll 8
        for (int k = 0; k < BS_DRAWS_PER_THREAD; ++k) {
            const int64_t j = j0 + (int64_t)k * BS_THREADS + threadIdx.x;
            if (j < sample_size) {
                const float v = shard_fetch(e, (uint32_t)row[j], shard_lo, shard_len);
                if ((k & 3) == 0) acc0 += v; else if ((k & 3) == 1) acc1 += v; else if ((k & 3) == 2) acc2 += v; else acc3 += v;
            }
        }
    } else {
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll 4
        for (int k = 0; k < BS_DRAWS_PER_THREAD / 4; ++k) {
            // 4 consecutive draws of this resample per Philox call
            const int64_t j = j0 + ((int64_t)k * BS_THREADS + threadIdx.x) * 4;
            if (j < sample_size) {
                const uint64_t c = offset + ((uint64_t)r * (uint64_t)stride4) + (uint64_t)(j >> 2);
                const uint4 u = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
                acc0 += shard_fetch(e, __umulhi(u.x, max_idx), shard_lo, shard_len);
                if (j + 1 < sample_size) acc1 += shard_fetch(e, __umulhi(u.y, max_idx), shard_lo, shard_len);
                if (j + 2 < sample_size) acc2 += shard_fetch(e, __umulhi(u.z, max_idx), shard_lo, shard_len);
                if (j + 3 < sample_size) acc3 += shard_fetch(e, __umulhi(u.w, max_idx), shard_lo, shard_len);
            }
        }
    }
    double s = (double)acc0 + (double)acc1 + (double)acc2 + (double)acc3;
    s = warp_sum(s);
    __shared__ double ws[BS_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < BS_THREADS / 32 ? ws[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0) atomicAdd(out + r, s);
    }
}

// ---------------------------------------------------------------------------------------------
// Bayesian bootstrap (analysis/bootstrap.py:236-262 with statistic = fep_estimator, estimator.py:78-79):
// sample weights ~ Dirichlet(1, ..., 1) = normalised Exp(1) variables g_ri = -log(u_ri), so that
//     -kT logsumexp(v_i + log weight_ri) = -kT [ log sum_i e^{v_i} g_ri - log sum_i g_ri ].
// One streaming pass per resample over the exp table (coalesced, no batch x n weight matrix): the uniforms
// come from Philox4x32-10 in registers, counter = (resample, sample / 4).
// grid = (chunks of BS_CHUNK samples, resamples); out_s[r] += sum e_i g_ri, out_g[r] += sum g_ri.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float exp1_variate(uint32_t u) {
    // u uniform on {0, .., 2^32 - 1} -> (u + 0.5) 2^-32 in (0, 1); -log of it
    return -0.6931471805599453f * __log2f(((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f));
}

__global__ void __launch_bounds__(BS_THREADS) bayesian_sums_kernel(const float* __restrict__ e, int64_t n, uint64_t seed,
                                                                   uint64_t offset, double* __restrict__ out_s,
                                                                   double* __restrict__ out_g) {
    const int r = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * BS_CHUNK;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    float s0 = 0.f, s1 = 0.f, g0 = 0.f, g1 = 0.f;
    const bool vec = (reinterpret_cast<uintptr_t>(e) & 15) == 0;
#pragma unroll 2
    for (int k = 0; k < BS_DRAWS_PER_THREAD / 4; ++k) {
        const int64_t j = j0 + ((int64_t)k * BS_THREADS + threadIdx.x) * 4;      // 4 consecutive samples per Philox call
        if (j < n) {
            const uint64_t c = offset + ((uint64_t)r * (uint64_t)((n + 3) / 4)) + (uint64_t)(j >> 2);
            const uint4 u = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
            const float a = exp1_variate(u.x), b = exp1_variate(u.y), cc = exp1_variate(u.z), d = exp1_variate(u.w);
            if (vec && j + 3 < n) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(e + j));
                s0 = fmaf(v.x, a, s0); s1 = fmaf(v.y, b, s1); s0 = fmaf(v.z, cc, s0); s1 = fmaf(v.w, d, s1);
                g0 += a + cc; g1 += b + d;
            } else {
                s0 = fmaf(__ldg(e + j), a, s0); g0 += a;
                if (j + 1 < n) { s1 = fmaf(__ldg(e + j + 1), b, s1); g1 += b; }
                if (j + 2 < n) { s0 = fmaf(__ldg(e + j + 2), cc, s0); g0 += cc; }
                if (j + 3 < n) { s1 = fmaf(__ldg(e + j + 3), d, s1); g1 += d; }
            }
        }
    }
    double s = (double)s0 + (double)s1, g = (double)g0 + (double)g1;
    s = warp_sum(s);
    g = warp_sum(g);
    __shared__ double ws[BS_THREADS / 32], wg[BS_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wg[threadIdx.x >> 5] = g; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < BS_THREADS / 32 ? ws[threadIdx.x] : 0.0;
        g = threadIdx.x < BS_THREADS / 32 ? wg[threadIdx.x] : 0.0;
        s = warp_sum(s);
        g = warp_sum(g);
        if (threadIdx.x == 0) { atomicAdd(out_s + r, s); atomicAdd(out_g + r, g); }
    }
}

// ---------------------------------------------------------------------------------------------
// MT19937: one CTA advances the 624-word state by whole twists (three dependency-free spans of
// 227 words, see oracle/analysis_oracle.py) and tempers / reduces the outputs modulo max_idx.
// The stream is inherently sequential, so a single CTA is used; jump-ahead is future work.
// ---------------------------------------------------------------------------------------------
constexpr int MT_N = 624, MT_M = 397, MT_THREADS = 256;

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    return y ^ (y >> 18);
}

__global__ void __launch_bounds__(MT_THREADS) mt19937_indices_kernel(uint32_t* __restrict__ state625, int64_t count,
                                                                     uint32_t max_idx, int32_t* __restrict__ idx) {
    __shared__ uint32_t st[MT_N];
    const int t = threadIdx.x;
    for (int i = t; i < MT_N; i += MT_THREADS) st[i] = state625[i];
    int pos = (int)state625[MT_N];
    __syncthreads();
    int64_t done = 0;
    while (done < count) {
        if (pos >= MT_N) {
            constexpr int SPAN = MT_N - MT_M;   // 227
            uint32_t v = 0;
            // span A: i in [0, 227)
            if (t < SPAN) v = st[t + MT_M] ^ mt_mix(st[t], st[t + 1]);
            __syncthreads();
            if (t < SPAN) st[t] = v;
            __syncthreads();
            // span B: i in [227, 454)
            if (t < SPAN) v = st[t] ^ mt_mix(st[t + SPAN], st[t + SPAN + 1]);
            __syncthreads();
            if (t < SPAN) st[t + SPAN] = v;
            __syncthreads();
            // span C: i in [454, 623) and the last word, which needs the NEW st[0]
            const int i = t + 2 * SPAN;
            if (i < MT_N - 1) v = st[i - SPAN] ^ mt_mix(st[i], st[i + 1]);
            else if (i == MT_N - 1) v = st[MT_M - 1] ^ mt_mix(st[MT_N - 1], st[0]);
            __syncthreads();
            if (i < MT_N) st[i] = v;
            __syncthreads();
            pos = 0;
        }
        const int avail = MT_N - pos;
        const int take = (int)((count - done) < (int64_t)avail ? (count - done) : (int64_t)avail);
        for (int k = t; k < take; k += MT_THREADS) idx[done + k] = (int32_t)(mt_temper(st[pos + k]) % max_idx);
        pos += take;
        done += take;
    }
    __syncthreads();
    for (int i = t; i < MT_N; i += MT_THREADS) state625[i] = st[i];
    if (t == 0) state625[MT_N] = (uint32_t)pos;
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

extern "C" int64_t tfepb_lse_workspace_bytes(void) { return (int64_t)LSE_MAX_BLOCKS * 2 * sizeof(double); }

// first pass of tfepb_lse / tfepb_fep_estimate: per-block (max, sum) pairs into `partials`; returns the number of blocks (< 0: error)
static int launch_lse_partial(int32_t dtype, const void* w, const void* logw, int64_t n, double scale, void* partials, cudaStream_t s) {
    int64_t blocks = (n + (int64_t)LSE_THREADS * 16 - 1) / ((int64_t)LSE_THREADS * 16);
    // one wave of co-resident blocks (4 per SM by the launch bounds): no tail wave
    const int64_t cap = (int64_t)sm_count() * 4 < LSE_MAX_BLOCKS ? (int64_t)sm_count() * 4 : LSE_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (dtype == TFEPB_F32)
        lse_partial_kernel<float><<<(int)blocks, LSE_THREADS, 0, s>>>((const float*)w, (const float*)logw, n, (float)scale,
                                                                      (double*)partials);
    else if (dtype == TFEPB_F64)
        lse_partial_kernel<double><<<(int)blocks, LSE_THREADS, 0, s>>>((const double*)w, (const double*)logw, n, scale,
                                                                       (double*)partials);
    else
        return fail(-1, "unknown dtype %d", dtype);
    if (check_launch("lse_partial")) return -1;
    return (int)blocks;
}

extern "C" int tfepb_lse(int32_t dtype, const void* w, const void* logw, int64_t n, double scale, void* partials,
                         double* out2, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(n > 0, "empty data");
    TFEPB_CHECK_ARG(w && partials && out2, "null buffer");
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = as_stream(stream);
    const int blocks = launch_lse_partial(dtype, w, logw, n, scale, partials, s);
    if (blocks < 0) return -1;
    lse_final_kernel<<<1, LSE_THREADS, 0, s>>>((const double*)partials, blocks, out2);
    return check_launch("lse_final");
}

extern "C" int tfepb_fep_estimate(int32_t dtype, const void* w, const void* logw, int64_t n, double kT, double log_norm,
                                  void* partials, double* out2, void* result, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(n > 0, "empty data");
    TFEPB_CHECK_ARG(w && partials && out2 && result, "null buffer");
    TFEPB_CHECK_ARG(kT > 0.0, "kT must be positive");
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = as_stream(stream);
    const int blocks = launch_lse_partial(dtype, w, logw, n, -1.0 / kT, partials, s);
    if (blocks < 0) return -1;
    if (dtype == TFEPB_F32)
        lse_final_estimate_kernel<float><<<1, LSE_THREADS, 0, s>>>((const double*)partials, blocks, out2, -kT, log_norm, (float*)result);
    else
        lse_final_estimate_kernel<double><<<1, LSE_THREADS, 0, s>>>((const double*)partials, blocks, out2, -kT, log_norm, (double*)result);
    return check_launch("lse_final_estimate");
}

extern "C" int tfepb_exp_table(int32_t dtype, const void* w, int64_t n, double scale, const double* max_dev, float* e,
                               tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(n > 0, "empty data");
    TFEPB_CHECK_ARG(w && max_dev && e, "null buffer");
    if (int rc = require_sm100()) return rc;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == TFEPB_F32)
        exp_table_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>((const float*)w, n, (float)scale, max_dev, e);
    else if (dtype == TFEPB_F64)
        exp_table_kernel<double><<<(int)blocks, 256, 0, as_stream(stream)>>>((const double*)w, n, scale, max_dev, e);
    else
        return fail(-1, "unknown dtype %d", dtype);
    return check_launch("exp_table");
}

extern "C" int tfepb_bootstrap_sums(const float* e, int64_t n, int64_t shard_lo, uint32_t max_idx, const int32_t* idx,
                                    int64_t ldidx, int32_t n_resamples, int64_t sample_size, uint64_t philox_seed,
                                    uint64_t philox_offset, const int64_t* sample_sizes, double* out_sums,
                                    tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(e && out_sums, "null buffer");
    TFEPB_CHECK_ARG(sample_sizes == nullptr || idx == nullptr, "per-resample sizes go with the Philox stream only");
    TFEPB_CHECK_ARG(n_resamples > 0 && sample_size > 0, "bad sizes");
    TFEPB_CHECK_ARG(max_idx > 0 && shard_lo >= 0 && n > 0 && n <= 0x7fffffff, "index range out of bounds");
    TFEPB_CHECK_ARG(n_resamples <= 65535, "at most 65535 resamples per call");
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = as_stream(stream);
    TFEPB_CUDA(cudaMemsetAsync(out_sums, 0, sizeof(double) * n_resamples, s));
    dim3 grid((unsigned)((sample_size + BS_CHUNK - 1) / BS_CHUNK), (unsigned)n_resamples);
    bootstrap_sums_kernel<<<grid, BS_THREADS, 0, s>>>(e, max_idx, idx, ldidx, sample_size, philox_seed, philox_offset,
                                                      (uint32_t)shard_lo, (uint32_t)n, sample_sizes, out_sums);
    return check_launch("bootstrap_sums");
}

extern "C" int tfepb_bayesian_bootstrap_sums(const float* e, int64_t n, int32_t n_resamples, uint64_t philox_seed,
                                             uint64_t philox_offset, double* out_sums, double* out_weight_sums,
                                             tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(e && out_sums && out_weight_sums, "null buffer");
    TFEPB_CHECK_ARG(n > 0 && n_resamples > 0, "bad sizes");
    TFEPB_CHECK_ARG(n_resamples <= 65535, "at most 65535 resamples per call");
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = as_stream(stream);
    TFEPB_CUDA(cudaMemsetAsync(out_sums, 0, sizeof(double) * n_resamples, s));
    TFEPB_CUDA(cudaMemsetAsync(out_weight_sums, 0, sizeof(double) * n_resamples, s));
    dim3 grid((unsigned)((n + BS_CHUNK - 1) / BS_CHUNK), (unsigned)n_resamples);
    bayesian_sums_kernel<<<grid, BS_THREADS, 0, s>>>(e, n, philox_seed, philox_offset, out_sums, out_weight_sums);
    return check_launch("bayesian_bootstrap_sums");
}

extern "C" int tfepb_mt19937_seed(uint32_t seed, uint32_t* state625_host) {
    TFEPB_CHECK_ARG(state625_host != nullptr, "null buffer");
    state625_host[0] = seed;
    for (int i = 1; i < MT_N; ++i)
        state625_host[i] = 1812433253u * (state625_host[i - 1] ^ (state625_host[i - 1] >> 30)) + (uint32_t)i;
    state625_host[MT_N] = MT_N;   // exhausted: the first draw twists
    return 0;
}

extern "C" int tfepb_mt19937_indices(uint32_t* state625_dev, int64_t count, uint32_t max_idx, int32_t* idx,
                                     tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(state625_dev && idx, "null buffer");
    TFEPB_CHECK_ARG(count >= 0 && max_idx > 0, "bad sizes");
    TFEPB_CHECK_ARG(max_idx <= 0x7fffffffu, "indices are int32: max_idx must be below 2^31");
    if (int rc = require_sm100()) return rc;
    if (count == 0) return 0;
    mt19937_indices_kernel<<<1, MT_THREADS, 0, as_stream(stream)>>>(state625_dev, count, max_idx, idx);
    return check_launch("mt19937_indices");
}
