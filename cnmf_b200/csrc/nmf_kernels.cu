// Elementwise / reduction kernels of the batched NMF engine (see nmf_kernels.cuh for the layout).
// All of these are HBM-bound streaming kernels: coalesced along the item (cell / gene) axis,
// the tiny K x K Gram matrices are broadcast from shared memory, fp64 only for the scalars
// that feed the convergence decisions.
#include "nmf_kernels.cuh"

namespace cnmf {

namespace {

constexpr float EPSILON_F32 = 1.1920928955078125e-07f;   // np.finfo(np.float32).eps, sklearn _nmf.py:32

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  T r = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (warp == 0) r = warp_sum(r);
  return r;   // valid in thread 0
}

// ------------------------------------------------------------------ split / transpose / sums
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

__global__ void split_scaled_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                    int rows, int ld, const float* __restrict__ col_scale) {
  const int ld4 = ld / 4;
  const long long n4 = (long long)rows * ld4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % ld4);
    float4 v = reinterpret_cast<const float4*>(src)[i];
    if (col_scale) {
      const float4 sc = reinterpret_cast<const float4*>(col_scale)[c4];
      v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
    }
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// positive floats order like their bit patterns: atomicMin on the int view
__global__ void min_positive_kernel(const float* __restrict__ X, int rows, int cols, int ld, int* __restrict__ col_min,
                                    int* __restrict__ row_min) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * 64;
  int cm = 0x7f800000;
  for (int r = r0; r < min(rows, r0 + 64); ++r) {
    const float v = (c < cols) ? X[(long long)r * ld + c] : 0.f;
    int b = (v > 0.f) ? __float_as_int(v) : 0x7f800000;
    cm = min(cm, b);
    // row minimum: warp reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b != 0x7f800000) atomicMin(&row_min[r], b);
  }
  if (c < cols && cm != 0x7f800000) atomicMin(&col_min[c], cm);
}

__global__ void check_scaled_int_kernel(const float* __restrict__ X, int rows, int cols, int ld,
                                        const float* __restrict__ rs, const float* __restrict__ cs, int* __restrict__ n_bad) {
  int bad = 0;
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = X[(long long)r * ld + c];
    if (v == 0.f) continue;
    const float sc = (rs ? rs[r] : 1.f) * (cs ? cs[c] : 1.f);
    const float q = v / sc;
    const float n = rintf(q);
    if (!(v > 0.f) || !(n >= 1.f) || n > 2048.f || fabsf(q - n) > 1e-4f * n) ++bad;
  }
  bad = warp_sum(bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

__global__ void build_counts_kernel(const float* __restrict__ X, int rows, int cols, int ld, const float* __restrict__ rs,
                                    const float* __restrict__ cs, float* __restrict__ C) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = X[(long long)r * ld + c];
    const float sc = (rs ? rs[r] : 1.f) * (cs ? cs[c] : 1.f);
    C[(long long)r * ld + c] = (v == 0.f) ? 0.f : rintf(v / sc);
  }
}

__global__ void fix_scale_kernel(float* __restrict__ v, int n, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i >= n) { v[i] = 0.f; return; }
  const float x = v[i];
  v[i] = (isfinite(x) && x > 0.f && x < 1e30f) ? x : 1.f;   // rows / columns without a positive entry: any scale works
}

__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ dst,
                                 float* __restrict__ dst_hi, float* __restrict__ dst_lo, int ld_dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;      // dst row = c, dst col = r
    if (c < cols && r < rows) {
      const float v = tile[threadIdx.x][i];
      const long long o = (long long)c * ld_dst + r;
      if (dst) dst[o] = v;
      if (dst_hi) {
        float h, l;
        split_tf32(v, h, l);
        dst_hi[o] = h;
        dst_lo[o] = l;
      }
    }
  }
}

__global__ void sums_partial_kernel(const float* __restrict__ X, int rows, int cols, int ld, double* __restrict__ part) {
  __shared__ double sm[32];
  double s = 0, q = 0;
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const double v = X[(long long)r * ld + c];
    s += v;
    q += v * v;
  }
  s = block_sum(s, sm);
  q = block_sum(q, sm);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = s;
    part[2 * blockIdx.x + 1] = q;
  }
}
__global__ void sums_final_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out2) {
  __shared__ double sm[32];
  double s = 0, q = 0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
    s += part[2 * i];
    q += part[2 * i + 1];
  }
  s = block_sum(s, sm);
  q = block_sum(q, sm);
  if (threadIdx.x == 0) {
    out2[0] = s;
    out2[1] = q;
  }
}

// ------------------------------------------------------------------ update kernels
// Thread = one item (cell / gene) column of one restart.  The unrolled K x K work is instantiated at a
// granularity of 4 components (KP = K rounded up to a multiple of 4) and dispatched per restart INSIDE the
// kernel (block-uniform switch), so a K = 10 restart runs the 12 x 12 body even when the batch also holds
// K = 13 restarts.  KPMAX (16 or 32, from the batch maximum) only bounds which bodies exist, i.e. the
// kernel's register budget: 80 registers / 3 blocks per SM for KPMAX = 16.  The K x K Gram is re-read
// from shared memory for every column through a volatile pointer (broadcast LDS.128) instead of being
// hoisted into ~256 registers (that version ran at ~1 TB/s).
__device__ __forceinline__ void load_gram_smem(float* G, const GramRef& gram, int r, int K, int KP, float diag_add) {
  for (int idx = threadIdx.x; idx < KP * KP; idx += blockDim.x) {
    const int c = idx / KP, i = idx % KP;
    double a = 0.0;
    if (c < K && i < K) {
      const double* p = gram.part + (long long)r * gram.chunks * gram.stride + idx;
      for (int ch = 0; ch < gram.chunks; ++ch) a += p[(long long)ch * gram.stride];    // fixed order: deterministic
    }
    float g = (float)a;
    if (c == i && c < K) g += diag_add;
    G[idx] = g;
  }
  __syncthreads();
}

template <int KP>
__device__ __forceinline__ void load_column(const FactorView& f, const float* __restrict__ NUM, int nsplit,
                                            long long sstride, int K, int o, int col, float (&fv)[KP], float (&nv)[KP]) {
  // all loads of this column first (2K independent requests in flight per thread)
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const long long e = (long long)(o + i) * f.ld + col;
    fv[i] = (i < K) ? f.F[e] : 0.f;
    float num = (i < K) ? NUM[e] : 0.f;
    for (int s = 1; s < nsplit; ++s) num += (i < K) ? NUM[s * sstride + e] : 0.f;
    nv[i] = num;
  }
}

__device__ __forceinline__ void store_elem(const FactorView& f, long long e, float v, float pscale) {
  f.F[e] = v;
  if (f.F_hi) {
    float h, l;
    split_tf32(v * pscale, h, l);
    f.F_hi[e] = h;
    f.F_lo[e] = l;
  }
}

template <int KP>
__device__ __forceinline__ double mu_body(const FactorView& f, const float* __restrict__ NUM, int nsplit, long long sstride,
                                          const float* G, int K, int o, float l1, float l2, int col_begin, int col_end,
                                          bool want_cross) {
  const volatile float4* Gv = reinterpret_cast<const volatile float4*>(G);
  double cross = 0.0;
  for (int col = col_begin + threadIdx.x; col < col_end; col += UPD_THREADS) {
    float fv[KP], nv[KP];
    load_column<KP>(f, NUM, nsplit, sstride, K, o, col, fv, nv);
    const float pscale = f.piece_scale ? f.piece_scale[col] : 1.f;
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      if (c < K) {
        float den = 0.f;
#pragma unroll
        for (int i4 = 0; i4 < KP / 4; ++i4) {
          const volatile float4& gq = Gv[c * (KP / 4) + i4];
          den = fmaf(gq.x, fv[4 * i4 + 0], den);
          den = fmaf(gq.y, fv[4 * i4 + 1], den);
          den = fmaf(gq.z, fv[4 * i4 + 2], den);
          den = fmaf(gq.w, fv[4 * i4 + 3], den);
        }
        if (l1 > 0.f) den += l1;
        if (l2 > 0.f) den += l2 * fv[c];
        if (den == 0.f) den = EPSILON_F32;
        const float fn = fv[c] * (nv[c] / den);
        store_elem(f, (long long)(o + c) * f.ld + col, fn, pscale);
        if (want_cross) cross += (double)nv[c] * (double)fn;
      }
    }
  }
  return cross;
}

template <int KP>
__device__ __forceinline__ double cd_body(const FactorView& f, const float* __restrict__ NUM, int nsplit, long long sstride,
                                          const float* G, int K, int o, float l1, int col_begin, int col_end) {
  const volatile float4* Gv = reinterpret_cast<const volatile float4*>(G);
  const volatile float* Gs = G;
  double viol = 0.0;
  for (int col = col_begin + threadIdx.x; col < col_end; col += UPD_THREADS) {
    float fv[KP], nv[KP];
    load_column<KP>(f, NUM, nsplit, sstride, K, o, col, fv, nv);
    const float pscale = f.piece_scale ? f.piece_scale[col] : 1.f;
#pragma unroll
    for (int t = 0; t < KP; ++t) {
      if (t < K) {
        float g = l1 - nv[t];                           // -(XHt - l1), sklearn _nmf.py:386-388
#pragma unroll
        for (int i4 = 0; i4 < KP / 4; ++i4) {           // same summation order as the Cython loop (r = 0..K-1)
          const volatile float4& gq = Gv[t * (KP / 4) + i4];
          g = fmaf(gq.x, fv[4 * i4 + 0], g);
          g = fmaf(gq.y, fv[4 * i4 + 1], g);
          g = fmaf(gq.z, fv[4 * i4 + 2], g);
          g = fmaf(gq.w, fv[4 * i4 + 3], g);
        }
        const float pg = (fv[t] == 0.f) ? fminf(0.f, g) : g;
        viol += (double)fabsf(pg);
        const float h = Gs[t * KP + t];
        if (h != 0.f) fv[t] = fmaxf(fv[t] - g / h, 0.f);
        store_elem(f, (long long)(o + t) * f.ld + col, fv[t], pscale);
      }
    }
  }
  return viol;
}

#define CNMF_KP_SWITCH(K, KPMAX, CALL)                                            \
  switch (((K) + 3) / 4) {                                                        \
    case 1: { constexpr int KP = 4; CALL; } break;                                \
    case 2: { constexpr int KP = 8; CALL; } break;                                \
    case 3: { constexpr int KP = 12; CALL; } break;                               \
    case 4: { constexpr int KP = 16; CALL; } break;                               \
    case 5: if constexpr (KPMAX >= 20) { constexpr int KP = 20; CALL; } break;    \
    case 6: if constexpr (KPMAX >= 24) { constexpr int KP = 24; CALL; } break;    \
    case 7: if constexpr (KPMAX >= 28) { constexpr int KP = 28; CALL; } break;    \
    default: if constexpr (KPMAX >= 32) { constexpr int KP = 32; CALL; } break;   \
  }

template <int KPMAX>
__global__ void __launch_bounds__(UPD_THREADS, KPMAX == 32 ? 2 : 3)
mu_update_kernel(FactorView f, const float* __restrict__ NUM, int nsplit, long long sstride,
                 GramRef gram, BatchMeta b, float l1, float l2, double* __restrict__ cross_partial) {
  const int slot = blockIdx.y;
  const int r = b.rid[slot];
  if (b.done[r]) return;
  const int K = b.k[slot], o = b.off[slot];
  __shared__ __align__(16) float G[KPMAX * KPMAX];
  __shared__ double red[32];
  const int col_begin = blockIdx.x * f.cpb;
  const int col_end = min(f.n, col_begin + f.cpb);
  double cross = 0.0;
  CNMF_KP_SWITCH(K, KPMAX, (load_gram_smem(G, gram, r, K, KP, 0.f),
                            cross = mu_body<KP>(f, NUM, nsplit, sstride, G, K, o, l1, l2, col_begin, col_end, cross_partial != nullptr)));
  if (cross_partial) {
    cross = block_sum(cross, red);
    if (threadIdx.x == 0) cross_partial[(long long)r * gridDim.x + blockIdx.x] = cross;
  }
}

template <int KPMAX>
__global__ void __launch_bounds__(UPD_THREADS, KPMAX == 32 ? 2 : 3)
cd_update_kernel(FactorView f, const float* __restrict__ NUM, int nsplit, long long sstride,
                 GramRef gram, BatchMeta b, float l1, float l2, double* __restrict__ viol_partial) {
  const int slot = blockIdx.y;
  const int r = b.rid[slot];
  if (b.done[r]) return;
  const int K = b.k[slot], o = b.off[slot];
  __shared__ __align__(16) float G[KPMAX * KPMAX];
  __shared__ double red[32];
  const int col_begin = blockIdx.x * f.cpb;
  const int col_end = min(f.n, col_begin + f.cpb);
  double viol = 0.0;
  CNMF_KP_SWITCH(K, KPMAX, (load_gram_smem(G, gram, r, K, KP, l2),   // l2 on the diagonal: sklearn _nmf.py:383-385
                            viol = cd_body<KP>(f, NUM, nsplit, sstride, G, K, o, l1, col_begin, col_end)));
  if (viol_partial) {
    viol = block_sum(viol, red);
    if (threadIdx.x == 0) viol_partial[(long long)r * gridDim.x + blockIdx.x] = viol;
  }
}

// ------------------------------------------------------------------ <NUM, F> without update
__global__ void __launch_bounds__(UPD_THREADS)
cross_kernel(FactorView f, const float* __restrict__ NUM, int nsplit, long long sstride, BatchMeta b,
             double* __restrict__ cross_partial) {
  const int slot = blockIdx.y;
  const int r = b.rid[slot];
  if (b.done[r]) return;
  const int K = b.k[slot], o = b.off[slot];
  __shared__ double red[32];
  const int col_begin = blockIdx.x * f.cpb;
  const int col_end = min(f.n, col_begin + f.cpb);
  double cross = 0.0;
  for (int col = col_begin + threadIdx.x; col < col_end; col += UPD_THREADS) {
    for (int c = 0; c < K; ++c) {
      const long long e = (long long)(o + c) * f.ld + col;
      float num = NUM[e];
      for (int s = 1; s < nsplit; ++s) num += NUM[s * sstride + e];
      cross += (double)num * (double)f.F[e];
    }
  }
  cross = block_sum(cross, red);
  if (threadIdx.x == 0) cross_partial[(long long)r * gridDim.x + blockIdx.x] = cross;
}

// ------------------------------------------------------------------ K x K Gram partials
// Register-tiled: a thread owns RB rows x KP columns of the K x K Gram and walks over columns of F,
// VEC columns at a time through 8/16-byte loads (KP*VEC values in flight per thread; a first version with
// one 4-byte column per step was latency-bound at ~0.6 TB/s).  TPC threads cooperate on one column group;
// partial sums are fp32 over the thread's columns, then fp64 through shuffles + a fixed-order
// shared-memory reduction (deterministic).  Same per-restart KP dispatch as the update kernels.
template <int KP, int BT>
struct GramCfg {
  static constexpr int TPC = KP <= 8 ? 1 : (KP <= 16 ? 2 : (KP <= 24 ? 4 : 8));   // threads per column group
  static constexpr int RB = (KP + TPC - 1) / TPC;       // rows of the Gram per thread (last block may be partial)
  static constexpr int VEC = KP <= 12 ? 4 : 2;          // consecutive columns per load (register budget)
  static constexpr int COLS_PER_ITER = (BT / TPC) * VEC;
  static constexpr int WARPS = BT / 32;
};

template <int VEC> struct VecLoad;
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
    const float2 q = *reinterpret_cast<const float2*>(p);
    v[0] = q.x; v[1] = q.y;
  }
};

template <int KP, int BT>
__device__ __forceinline__ void gram_body(const FactorView& f, int K, int o, int col_begin, int col_end,
                                          double* part /* smem WARPS x 8 x 32 */, double* __restrict__ out) {
  using C = GramCfg<KP, BT>;
  const int rb = threadIdx.x % C::TPC;                 // which row block of the Gram
  const int cl = threadIdx.x / C::TPC;                 // column-group lane inside the block
  float acc[C::RB][KP];
#pragma unroll
  for (int a = 0; a < C::RB; ++a)
#pragma unroll
    for (int i = 0; i < KP; ++i) acc[a][i] = 0.f;
  const float* __restrict__ Fp = f.F;
  for (int col = col_begin + cl * C::VEC; col < col_end; col += C::COLS_PER_ITER) {   // padding columns hold zeros
    float fv[KP][C::VEC];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      if (i < K) {
        VecLoad<C::VEC>::ld(Fp + (long long)(o + i) * f.ld + col, fv[i]);
      } else {
#pragma unroll
        for (int u = 0; u < C::VEC; ++u) fv[i][u] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < C::VEC; ++u) {
#pragma unroll
      for (int a = 0; a < C::RB; ++a) {
        float fa = 0.f;                                 // fv[rb * RB + a][u] without dynamic register indexing
#pragma unroll
        for (int t = 0; t < C::TPC; ++t)
          if (t == rb && t * C::RB + a < KP) fa = fv[t * C::RB + a < KP ? t * C::RB + a : 0][u];
#pragma unroll
        for (int i = 0; i < KP; ++i) acc[a][i] = fmaf(fa, fv[i][u], acc[a][i]);
      }
    }
  }
  // reduction over the column lanes: xor-shuffles among the lanes that share a row block (lane % TPC),
  // then the 8 warps' partials are summed in fixed order through shared memory -- all in fp64
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < C::RB; ++a) {
    double v[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) v[i] = (double)acc[a][i];
#pragma unroll
    for (int sh = 16; sh >= C::TPC; sh >>= 1)
#pragma unroll
      for (int i = 0; i < KP; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], sh);
    __syncthreads();
    if (lane < C::TPC) {
#pragma unroll
      for (int i = 0; i < KP; ++i) part[(warp * 8 + lane) * 32 + i] = v[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < C::TPC * KP; t += BT) {
      const int rbb = t / KP, i = t % KP;
      const int row = rbb * C::RB + a;
      if (row < KP) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) sum += part[(w * 8 + rbb) * 32 + i];
        out[row * KP + i] = sum;
      }
    }
  }
}

// BT = 256: stand-alone launches (one block per SM by registers).  BT = 32: one-warp blocks that fit beside a
// resident GEMM CTA (8 K registers, 2 KB smem), used when the Gram runs on the auxiliary stream under a GEMM.
template <int KPMAX, int BT>
__global__ void __launch_bounds__(BT)
gram_partial_kernel(FactorView f, BatchMeta b, double* __restrict__ gram_partial) {
  const int slot = blockIdx.y;
  const int r = b.rid[slot];
  if (b.done[r]) return;
  const int K = b.k[slot], o = b.off[slot];
  __shared__ double part[(BT / 32) * 8 * 32];
  const int col_begin = blockIdx.x * f.gcpb;
  const int col_end = min(f.n, col_begin + f.gcpb);
  double* out = gram_partial + ((long long)r * gridDim.x + blockIdx.x) * (KPMAX * KPMAX);
  CNMF_KP_SWITCH(K, KPMAX, (gram_body<KP, BT>(f, K, o, col_begin, col_end, part, out)));
}

__global__ void finalize_kernel(const double* __restrict__ gram_partial, double* __restrict__ gram,
                                const double* __restrict__ scal_partial, double* __restrict__ scal, int chunks,
                                BatchMeta b) {
  const int r = b.rid[blockIdx.x];
  if (b.done[r]) return;
  if (gram_partial) {
    const int K = b.k[blockIdx.x];
    const int KP = ((K + 3) / 4) * 4;                   // layout written by gram_body<KP>
    const int stride = b.kp * b.kp;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
      double a = 0.0;
      for (int ch = 0; ch < chunks; ++ch) a += gram_partial[((long long)r * chunks + ch) * stride + e];
      const int c = e / KP, i = e % KP;
      gram[(long long)r * KMAX * KMAX + c * KMAX + i] = a;
    }
  }
  if (scal_partial && threadIdx.x == 0) {
    double a = 0.0;
    for (int ch = 0; ch < chunks; ++ch) a += scal_partial[(long long)r * chunks + ch];
    scal[r] = a;
  }
}

// ------------------------------------------------------------------ convergence
__global__ void mu_check_kernel(ConvState st, const double* __restrict__ cross, const double* __restrict__ gramA,
                                const double* __restrict__ gramB, double normX2, BatchMeta b, int it, double tol,
                                int max_iter) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= b.R) return;
  const int r = b.rid[slot];
  if (st.done[r]) return;
  const int K = b.k[slot];
  double dot = 0.0;
  for (int c = 0; c < K; ++c)
    for (int i = 0; i < K; ++i)
      dot += gramA[(long long)r * KMAX * KMAX + c * KMAX + i] * gramB[(long long)r * KMAX * KMAX + c * KMAX + i];
  const double err = sqrt(fmax(normX2 - 2.0 * cross[r] + dot, 0.0));
  st.last[r] = err;
  if (it == 0) {
    st.err0[r] = err;
    st.prev[r] = err;
    return;
  }
  if ((st.prev[r] - err) / st.err0[r] < tol) {
    st.done[r] = 1;
    st.n_iter[r] = it;
  } else {
    st.prev[r] = err;
    if (it >= max_iter) {
      st.done[r] = 1;
      st.n_iter[r] = it;
    }
  }
}

__global__ void cd_check_kernel(ConvState st, const double* __restrict__ violA, const double* __restrict__ violB,
                                BatchMeta b, int it, double tol, int max_iter) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= b.R) return;
  const int r = b.rid[slot];
  if (st.done[r]) return;
  const double viol = violA[r] + (violB ? violB[r] : 0.0);
  st.last[r] = viol;
  if (it == 1) st.err0[r] = viol;
  const double v0 = st.err0[r];
  if (v0 == 0.0 || viol / v0 <= tol || it >= max_iter) {
    st.done[r] = 1;
    st.n_iter[r] = it;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ src_off,
                                   float* __restrict__ dst, const int* __restrict__ dst_off,
                                   const int* __restrict__ k, int ld) {
  const int r = blockIdx.x / KMAX, c = blockIdx.x % KMAX;
  if (c >= k[r]) return;
  const float4* s = reinterpret_cast<const float4*>(src + (long long)(src_off[r] + c) * ld);
  float4* d = reinterpret_cast<float4*>(dst + (long long)(dst_off[r] + c) * ld);
  for (int i = threadIdx.x; i < ld / 4; i += blockDim.x) d[i] = s[i];
}

}  // namespace

// ============================================================================ launchers
int launch_split_tf32(const float* src, float* hi, float* lo, long long n_elems, cudaStream_t s) {
  CNMF_REQUIRE(n_elems % 4 == 0, "split_tf32: element count must be a multiple of 4");
  const long long n4 = n_elems / 4;
  if (n4 == 0) return 0;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split_tf32_kernel<<<blocks, 256, 0, s>>>(src, hi, lo, n4);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_split_scaled(const float* src, float* hi, float* lo, int rows, int ld, const float* col_scale, cudaStream_t s) {
  CNMF_REQUIRE(ld % 4 == 0, "split_scaled: ld must be a multiple of 4");
  const long long n4 = (long long)rows * (ld / 4);
  if (n4 == 0) return 0;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split_scaled_kernel<<<blocks, 256, 0, s>>>(src, hi, lo, rows, ld, col_scale);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_min_positive(const float* X, int rows, int cols, int ld, float* col_min, float* row_min, cudaStream_t s) {
  CNMF_CUDA_CHECK(cudaMemsetAsync(col_min, 0x7f, sizeof(float) * cols, s));   // 0x7f7f7f7f: a huge finite float
  CNMF_CUDA_CHECK(cudaMemsetAsync(row_min, 0x7f, sizeof(float) * rows, s));
  dim3 grid((cols + 255) / 256, (rows + 63) / 64);
  CNMF_REQUIRE(grid.y <= 65535, "min_positive: too many rows for one launch");
  min_positive_kernel<<<grid, 256, 0, s>>>(X, rows, cols, ld, reinterpret_cast<int*>(col_min), reinterpret_cast<int*>(row_min));
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_check_scaled_int(const float* X, int rows, int cols, int ld, const float* row_scale, const float* col_scale,
                            int* n_bad, cudaStream_t s) {
  check_scaled_int_kernel<<<148 * 8, 256, 0, s>>>(X, rows, cols, ld, row_scale, col_scale, n_bad);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_build_counts(const float* X, int rows, int cols, int ld, const float* row_scale, const float* col_scale,
                        float* C, cudaStream_t s) {
  build_counts_kernel<<<148 * 8, 256, 0, s>>>(X, rows, cols, ld, row_scale, col_scale, C);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_fix_scale(float* v, int n, int n_pad, cudaStream_t s) {
  fix_scale_kernel<<<(n_pad + 255) / 256, 256, 0, s>>>(v, n, n_pad);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_transpose(const float* src, int rows, int cols, int ld_src, float* dst, float* dst_hi, float* dst_lo,
                     int ld_dst, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  CNMF_REQUIRE(grid.y <= 65535, "transpose: too many rows for one launch");
  transpose_kernel<<<grid, block, 0, s>>>(src, rows, cols, ld_src, dst, dst_hi, dst_lo, ld_dst);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_matrix_sums(const float* X, int rows, int cols, int ld, double* out2, double* scratch, int scratch_len,
                       cudaStream_t s) {
  int blocks = 148 * 8;
  if (2 * blocks > scratch_len) blocks = scratch_len / 2;
  CNMF_REQUIRE(blocks >= 1, "matrix_sums: scratch too small");
  sums_partial_kernel<<<blocks, 256, 0, s>>>(X, rows, cols, ld, scratch);
  sums_final_kernel<<<1, 256, 0, s>>>(scratch, blocks, out2);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

#define CNMF_DISPATCH_KPMAX(kp, CALL)                               \
  switch (kp) {                                                      \
    case 16: { constexpr int KPMAX = 16; CALL; } break;              \
    case 32: { constexpr int KPMAX = 32; CALL; } break;              \
    default: set_last_error("kp must be 16 or 32"); return -1;       \
  }

int launch_mu_update(const FactorView& f, const float* NUM, int nsplit, long long sstride, const GramRef& gram,
                     const BatchMeta& b, float l1, float l2, double* cross_partial, cudaStream_t s) {
  dim3 grid(col_chunks(f), b.R);
  CNMF_DISPATCH_KPMAX(b.kp, (mu_update_kernel<KPMAX><<<grid, UPD_THREADS, 0, s>>>(f, NUM, nsplit, sstride, gram, b, l1,
                                                                                    l2, cross_partial)));
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_cd_update(const FactorView& f, const float* NUM, int nsplit, long long sstride, const GramRef& gram,
                     const BatchMeta& b, float l1, float l2, double* viol_partial, cudaStream_t s) {
  dim3 grid(col_chunks(f), b.R);
  CNMF_DISPATCH_KPMAX(b.kp, (cd_update_kernel<KPMAX><<<grid, UPD_THREADS, 0, s>>>(f, NUM, nsplit, sstride, gram, b, l1,
                                                                                    l2, viol_partial)));
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_cross(const FactorView& f, const float* NUM, int nsplit, long long sstride, const BatchMeta& b,
                 double* cross_partial, cudaStream_t s) {
  dim3 grid(col_chunks(f), b.R);
  cross_kernel<<<grid, UPD_THREADS, 0, s>>>(f, NUM, nsplit, sstride, b, cross_partial);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_gram_partial(const FactorView& f, const BatchMeta& b, double* gram_partial, cudaStream_t s, bool one_warp_blocks) {
  dim3 grid(gram_chunks(f), b.R);
  if (one_warp_blocks) {
    CNMF_DISPATCH_KPMAX(b.kp, (gram_partial_kernel<KPMAX, 32><<<grid, 32, 0, s>>>(f, b, gram_partial)));
  } else {
    CNMF_DISPATCH_KPMAX(b.kp, (gram_partial_kernel<KPMAX, 256><<<grid, 256, 0, s>>>(f, b, gram_partial)));
  }
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_finalize(const double* gram_partial, double* gram, const double* scal_partial, double* scal, int chunks,
                    const BatchMeta& b, cudaStream_t s) {
  finalize_kernel<<<b.R, 256, 0, s>>>(gram_partial, gram, scal_partial, scal, chunks, b);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_mu_check(const ConvState& st, const double* cross, const double* gramA, const double* gramB, double normX2,
                    const BatchMeta& b, int it, double tol, int max_iter, cudaStream_t s) {
  mu_check_kernel<<<(b.R + 127) / 128, 128, 0, s>>>(st, cross, gramA, gramB, normX2, b, it, tol, max_iter);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_cd_check(const ConvState& st, const double* violA, const double* violB, const BatchMeta& b, int it,
                    double tol, int max_iter, cudaStream_t s) {
  cd_check_kernel<<<(b.R + 127) / 128, 128, 0, s>>>(st, violA, violB, b, it, tol, max_iter);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_gather_rows(const float* src, const int* src_off, float* dst, const int* dst_off, const int* k, int R,
                       int ld, cudaStream_t s) {
  if (R == 0) return 0;
  gather_rows_kernel<<<R * KMAX, 128, 0, s>>>(src, src_off, dst, dst_off, k, ld);
  CNMF_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cnmf
