// Stage 2 of the two-stage reduction: symmetric band (half-width NB) -> tridiagonal, VALUES ONLY, by Householder
// bulge chasing (Murata-Horikoshi / Lang), batched over the T matrices of a launch.  The transformations never
// touch row/column 0, so e_0 is preserved and the loss of interpolatory.py:124-126 only needs the resulting
// (d, e) together with beta0 (gamma.cu); nothing that needs vectors goes through this stage (bandsolve.cu solves
// in band coordinates), so no reflector is stored or applied to anything but the band itself.
// NumPy model: tests/twostage_model.py::stage2.
//
// Parallelisation: sweep j (annihilate column j below the sub-diagonal, chase the bulge to the bottom) is owned
// by one CTA (or warp); the CTAs of a matrix run as a software pipeline: hop h of sweep j+1 may start once sweep j
// has finished hop h+1 (progress counters in global memory, release/acquire).  Sweeps are handed out by a
// per-matrix TICKET counter (atomicAdd): the CTA that draws sweep j knows that every sweep < j is owned by a CTA
// that is already running, so the wait chain always bottoms out at a running CTA -- no co-residency requirement,
// no cooperative launch, any grid size, and the kernel can share the GPU with the GEMMs of another stream.  The
// working band (2 NB sub-diagonals, 4.3 MB per matrix at n = 8192) stays in L2.
#include <stdlib.h>

#include "internal.cuh"

namespace chd {

constexpr int CB = CHD_BAND;
constexpr int CLD = CHD_CHASE_LD;
constexpr int CS_ = CB + 1;               // shared-memory stride of the dense blocks
constexpr int CH_WARPS = 4;
constexpr int CH_DONE = 1 << 30;

struct ChaseArgs {
  double* bandW;   // T x n x CLD, bandW[i][d] = A(i, i-d)
  int* prog;       // T x n   hops finished by sweep j
  int* ticket;     // T       next sweep to hand out
  int* err;        // 1       set when a wait on the predecessor sweep timed out (results become NaN)
  int n;
};
constexpr unsigned long long CH_SPIN_LIMIT = 1ull << 26;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Householder from the warp-distributed vector x (lane l < len holds x[l]); returns v (v[0] = 1), tau, beta.
__device__ __forceinline__ void warp_reflector(double x, int lane, int len, double& v, double& tau, double& beta) {
  const double alpha = __shfl_sync(0xffffffffu, x, 0);
  double ss = (lane >= 1 && lane < len) ? x * x : 0.0;
  ss = warp_sum(ss);
  if (ss == 0.0) {
    beta = alpha; tau = 0.0;
    v = (lane == 0) ? 1.0 : 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + ss), alpha);
    tau = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    v = (lane == 0) ? 1.0 : ((lane < len) ? x * scale : 0.0);
  }
}

__global__ void __launch_bounds__(CH_WARPS * 32) chase_kernel(ChaseArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.y, n = a.n;
  double* bw = a.bandW + (size_t)t * n * CLD;
  int* prog = a.prog + (size_t)t * n;
  double* D = smem + (size_t)wid * (2 * CB * CS_ + 3 * CB);   // CB x CS_
  double* S = D + CB * CS_;                                   // CB x CS_
  double* vs = S + CB * CS_;                                  // CB
  double* ws = vs + CB;                                       // CB
  double* v2s = ws + CB;                                      // CB

  while (true) {
    int j = 0;
    if (lane == 0) j = atomicAdd(a.ticket + t, 1);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j > n - 3) break;
    int lo = j + 1, hi = min(j + CB, n - 1);
    int len = hi - lo + 1;
    int h = 0;
    bool dead = false;
    auto wait_prev = [&](int need) {
      int ok = 1;
      if (j > 0 && lane == 0) {
        unsigned long long spins = 0;
        while (ld_acquire(prog + j - 1) < need) {
          __nanosleep(40);
          if (++spins > CH_SPIN_LIMIT) { ok = 0; atomicExch(a.err, 1); break; }
        }
      }
      dead = !__shfl_sync(0xffffffffu, ok, 0);
    };
    wait_prev(2);
    if (dead) { if (lane == 0) st_release(prog + j, CH_DONE); break; }
    // first reflector of the sweep: column j, rows lo..hi
    double v, tau, beta;
    {
      const double x = (lane < len) ? __ldcg(bw + (size_t)(lo + lane) * CLD + (lane + 1)) : 0.0;
      warp_reflector(x, lane, len, v, tau, beta);
      if (lane < len) __stcg(bw + (size_t)(lo + lane) * CLD + (lane + 1), (lane == 0) ? beta : 0.0);
    }
    while (true) {
      // ---------------- all loads of the hop up front (64 independent requests per lane in flight)
      const int lo2 = hi + 1, hi2 = min(hi + CB, n - 1);
      const int len2 = hi2 - lo2 + 1;
      const int off = lo2 - lo;   // == len
      double dr[CB], sr[CB];
#pragma unroll
      for (int l = 0; l < CB; ++l)   // row l of D, lanes = d (column l - d)
        dr[l] = (l < len && lane <= l) ? __ldcg(bw + (size_t)(lo + l) * CLD + lane) : 0.0;
#pragma unroll
      for (int l = 0; l < CB; ++l)   // row l of S, lanes = column c: A(lo2+l, lo+c) at d = off + l - c
        sr[l] = (l < len2 && lane < len) ? __ldcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane)) : 0.0;
      // ---------------- diagonal block D = A[lo..hi, lo..hi]  <- H D H
#pragma unroll
      for (int l = 0; l < CB; ++l) {
        if (lane <= l) {
          D[l * CS_ + (l - lane)] = dr[l];
          D[(l - lane) * CS_ + l] = dr[l];
        }
      }
#pragma unroll
      for (int l = 0; l < CB; ++l) S[l * CS_ + lane] = sr[l];
      vs[lane] = v;
      __syncwarp();
      double p = 0.0;
#pragma unroll 8
      for (int c = 0; c < CB; ++c) p = fma(D[lane * CS_ + c], vs[c], p);
      p *= tau;
      const double al = -0.5 * tau * warp_sum(p * v);
      const double wv = p + al * v;
      ws[lane] = wv;
      __syncwarp();
#pragma unroll 8
      for (int c = 0; c < CB; ++c) D[lane * CS_ + c] -= v * ws[c] + wv * vs[c];
      __syncwarp();
      for (int l = 0; l < len; ++l)
        if (lane <= l) __stcg(bw + (size_t)(lo + l) * CLD + lane, D[l * CS_ + (l - lane)]);
      // ---------------- block below: S = A[lo2..hi2, lo..hi]  <- S H, then the next reflector from its first column
      bool last = (len2 <= 0);
      double v2 = 0.0, tau2 = 0.0;
      if (!last) {
        double u = 0.0;
#pragma unroll 8
        for (int c = 0; c < CB; ++c) u = fma(S[lane * CS_ + c], vs[c], u);
        u *= tau;
#pragma unroll 8
        for (int c = 0; c < CB; ++c) S[lane * CS_ + c] -= u * vs[c];
        __syncwarp();
        if (len2 >= 2) {
          double beta2;
          const double x = S[lane * CS_ + 0];
          warp_reflector(x, lane, len2, v2, tau2, beta2);
          v2s[lane] = v2;
          S[lane * CS_ + 0] = (lane == 0) ? beta2 : 0.0;
          __syncwarp();
          // z[c] = tau2 sum_l v2[l] S[l][c]  (lane = column c >= 1), S[l][c] -= v2[l] z[c]
          if (lane >= 1) {
            double z = 0.0;
#pragma unroll 8
            for (int l = 0; l < CB; ++l) z = fma(v2s[l], S[l * CS_ + lane], z);
            z *= tau2;
#pragma unroll 8
            for (int l = 0; l < CB; ++l) S[l * CS_ + lane] -= v2s[l] * z;
          }
          __syncwarp();
        } else {
          last = true;
        }
        for (int l = 0; l < len2; ++l)
          if (lane < len) __stcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane), S[l * CS_ + lane]);
      }
      // ---------------- publish progress, move on
      __syncwarp();
      ++h;
      if (lane == 0) {
        __threadfence();
        st_release(prog + j, last ? CH_DONE : h);
      }
      if (last) break;
      v = v2; tau = tau2; lo = lo2; hi = hi2; len = len2;
      wait_prev(h + 2);
      if (dead) { if (lane == 0) st_release(prog + j, CH_DONE); break; }
    }
    if (dead) break;
  }
}

// ---- CTA-per-sweep variant (default): the same pipeline, but a sweep is owned by a 4-warp CTA (thread = (row, one
// of four 8-column slices) of the 32 x 32 blocks), which cuts the serial part of a hop -- the dense updates -- by ~4.
// The lag between consecutive sweeps (2 hops + flag visibility) is what bounds the whole stage.
constexpr int CC_THREADS = 128;
constexpr int CC_PARTS = CC_THREADS / 32;
constexpr int CC_W = CB / CC_PARTS;   // columns (or rows) per slice
constexpr int CC_SMEM_DOUBLES = 2 * CB * CS_ + 4 * CB + CC_PARTS * CB;

__global__ void __launch_bounds__(CC_THREADS, 10) chase_cta_kernel(ChaseArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, part = tid >> 5;
  const int t = blockIdx.y, n = a.n;
  double* bw = a.bandW + (size_t)t * n * CLD;
  int* prog = a.prog + (size_t)t * n;
  __shared__ int s_j, s_ok, s_seen;
  double* D = smem;                    // CB x CS_
  double* S = D + CB * CS_;            // CB x CS_
  double* vs = S + CB * CS_;           // CB
  double* ws = vs + CB;                // CB
  double* v2s = ws + CB;               // CB
  double* zs = v2s + CB;               // CB
  double* red = zs + CB;               // CC_PARTS x CB
  const int c0 = part * CC_W;          // this thread's column slice (row = lane) / row slice (column = lane)

  while (true) {
    if (tid == 0) { s_j = atomicAdd(a.ticket + t, 1); s_ok = 1; }
    __syncthreads();
    const int j = s_j;
    if (j > n - 3) break;
    int lo = j + 1, hi = min(j + CB, n - 1);
    int len = hi - lo + 1;
    int h = 0;
    // returns false when the predecessor never got there (a bug or a dead context): flag it and let the host see
    // NaN results instead of trapping the whole context
    // `seen` = the predecessor's progress at the last look (uniform over the CTA): as long as it covers the hop we
    // are about to start, neither the acquire load (an L2 round trip on the critical path of every hop) nor the
    // barrier that broadcasts it is needed -- the release that published `seen` already ordered everything before it
    int seen = (j == 0) ? CH_DONE : 0;
    auto wait_prev = [&](int need) -> bool {
      if (seen >= need) return true;
      if (tid == 0) {
        unsigned long long spins = 0;
        int v;
        while ((v = ld_acquire(prog + j - 1)) < need)
          if (++spins > CH_SPIN_LIMIT) { s_ok = 0; atomicExch(a.err, 1); break; }
        s_seen = v;
      }
      __syncthreads();
      seen = s_seen;
      return s_ok != 0;
    };
    bool alive = wait_prev(2);
    if (!alive) { if (tid == 0) st_release(prog + j, CH_DONE); break; }
    // first reflector of the sweep: column j, rows lo..hi (every warp computes it: lane = row)
    double v, tau, beta;
    {
      const double x = (lane < len) ? __ldcg(bw + (size_t)(lo + lane) * CLD + (lane + 1)) : 0.0;
      warp_reflector(x, lane, len, v, tau, beta);
      if (part == 0 && lane < len) __stcg(bw + (size_t)(lo + lane) * CLD + (lane + 1), (lane == 0) ? beta : 0.0);
    }
    while (true) {
      const int lo2 = hi + 1, hi2 = min(hi + CB, n - 1);
      const int len2 = hi2 - lo2 + 1;
      const int off = lo2 - lo;
      // ---------------- loads: warp `part` fetches rows c0 .. c0 + CC_W - 1 of both blocks
      double dr[CC_W], sr[CC_W];
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        dr[i] = (l < len && lane <= l) ? __ldcg(bw + (size_t)(lo + l) * CLD + lane) : 0.0;
        sr[i] = (l < len2 && lane < len) ? __ldcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane)) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        if (lane <= l) {
          D[l * CS_ + (l - lane)] = dr[i];
          D[(l - lane) * CS_ + l] = dr[i];
        }
        S[l * CS_ + lane] = sr[i];
      }
      if (part == 0) vs[lane] = v;
      __syncthreads();
      // ---------------- D <- H D H   (thread: row = lane, columns c0 .. c0 + CC_W - 1)
      {
        double pp = 0.0;
#pragma unroll
        for (int c = 0; c < CC_W; ++c) pp = fma(D[lane * CS_ + c0 + c], vs[c0 + c], pp);
        red[part * CB + lane] = pp;
      }
      __syncthreads();
      double p = 0.0;
#pragma unroll
      for (int q = 0; q < CC_PARTS; ++q) p += red[q * CB + lane];
      p *= tau;
      const double al = -0.5 * tau * warp_sum(p * v);
      const double wv = p + al * v;
      if (part == 0) ws[lane] = wv;
      __syncthreads();
#pragma unroll
      for (int c = 0; c < CC_W; ++c) D[lane * CS_ + c0 + c] -= v * ws[c0 + c] + wv * vs[c0 + c];
      // ---------------- S <- S H
      bool last = (len2 <= 0);
      double v2 = 0.0, tau2 = 0.0;
      {
        double uu = 0.0;
#pragma unroll
        for (int c = 0; c < CC_W; ++c) uu = fma(S[lane * CS_ + c0 + c], vs[c0 + c], uu);
        __syncthreads();                     // red[] of the D part consumed, D slices written
        red[part * CB + lane] = uu;
      }
      // store D (rows c0 .. of this warp) while the S part goes on
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        if (l < len && lane <= l) __stcg(bw + (size_t)(lo + l) * CLD + lane, D[l * CS_ + (l - lane)]);
      }
      __syncthreads();
      if (!last) {
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < CC_PARTS; ++q) u += red[q * CB + lane];
        u *= tau;
#pragma unroll
        for (int c = 0; c < CC_W; ++c) S[lane * CS_ + c0 + c] -= u * vs[c0 + c];
        __syncthreads();
        if (len2 >= 2) {
          double beta2;
          const double x = S[lane * CS_ + 0];
          warp_reflector(x, lane, len2, v2, tau2, beta2);   // identical in every warp
          __syncthreads();                   // everybody has read column 0
          if (part == 0) {
            v2s[lane] = v2;
            S[lane * CS_ + 0] = (lane == 0) ? beta2 : 0.0;
          }
          __syncthreads();
          // z[c] = tau2 sum_l v2[l] S[l][c]: thread (column = lane, rows c0 .. c0 + CC_W - 1)
          {
            double zz = 0.0;
#pragma unroll
            for (int i = 0; i < CC_W; ++i) zz = fma(v2s[c0 + i], S[(c0 + i) * CS_ + lane], zz);
            red[part * CB + lane] = zz;
          }
          __syncthreads();
          double z = 0.0;
#pragma unroll
          for (int q = 0; q < CC_PARTS; ++q) z += red[q * CB + lane];
          z = (lane >= 1) ? z * tau2 : 0.0;
#pragma unroll
          for (int i = 0; i < CC_W; ++i) S[(c0 + i) * CS_ + lane] -= v2s[c0 + i] * z;
        } else {
          last = true;
        }
        // rows c0 .. of S were last written by this warp (left-apply) or, without it, by all warps (right-apply)
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CC_W; ++i) {
          const int l = c0 + i;
          if (l < len2 && lane < len) __stcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane), S[l * CS_ + lane]);
        }
      }
      // ---------------- publish progress, move on
      __syncthreads();
      ++h;
      if (tid == 0) {
        __threadfence();     // cumulative: the CTA's stores (ordered by the barrier) become visible before the flag
        st_release(prog + j, last ? CH_DONE : h);
      }
      if (last) break;
      v = v2; tau = tau2; lo = lo2; hi = hi2; len = len2;
      alive = wait_prev(h + 2);
      if (!alive) { if (tid == 0) st_release(prog + j, CH_DONE); break; }
    }
    if (!alive) break;
    __syncthreads();   // s_j is rewritten by thread 0 at the top of the loop
  }
}

// ---- CTA-per-sweep kernel, second formulation (default): the same arithmetic as chase_cta_kernel with 5 block barriers
// per hop instead of 11, no memory fence besides the release store, and the progress flag of hop h published after
// the first barrier of hop h + 1 (its stores are then ordered by that barrier) -- ncu showed the first formulation
// stalled on barriers (8.8 warps per issued instruction) and on the fence, not on arithmetic:
//   #1 blocks staged in shared memory      #2 partial products D v and S v (both at once)
//   #3 w of the two-sided update           #4 D, S H updated, next reflector from column 0 (every warp computes it from
//                                             the pre-update column, so nobody waits for the owner of column 0)
//   #5 partial products v2^T S;  the left update and the stores of S then run from registers.
constexpr int C2_SMEM_DOUBLES = 2 * CB * CS_ + 5 * CB + 2 * CC_PARTS * CB;

__global__ void __launch_bounds__(CC_THREADS, 10) chase_cta2_kernel(ChaseArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, part = tid >> 5;
  const int t = blockIdx.y, n = a.n;
  double* bw = a.bandW + (size_t)t * n * CLD;
  int* prog = a.prog + (size_t)t * n;
  __shared__ int s_j, s_ok, s_seen;
  double* D = smem;                    // CB x CS_
  double* S = D + CB * CS_;            // CB x CS_
  double* vs = S + CB * CS_;           // CB
  double* ws = vs + CB;                // CB
  double* v2s = ws + CB;               // CB
  double* col0 = v2s + CB;             // CB   column 0 of S after both updates
  double* spare = col0 + CB;           // CB
  double* red = spare + CB;            // CC_PARTS x CB
  double* red2 = red + CC_PARTS * CB;  // CC_PARTS x CB
  (void)spare;
  const int c0 = part * CC_W;          // this thread's column slice (row = lane) / row slice (column = lane)

  while (true) {
    __syncthreads();                   // s_j of the previous sweep has been read by everybody
    if (tid == 0) { s_j = atomicAdd(a.ticket + t, 1); s_ok = 1; }
    __syncthreads();
    const int j = s_j;
    if (j > n - 3) break;
    int lo = j + 1, hi = min(j + CB, n - 1);
    int len = hi - lo + 1;
    int h = 0;
    int pending = 0;                   // hops finished but not yet published
    int seen = (j == 0) ? CH_DONE : 0;
    auto wait_prev = [&](int need) -> bool {
      if (seen >= need) return true;
      __syncthreads();                 // the stores of the finished hops are ordered before the flag below
      if (tid == 0) {
        if (pending) st_release(prog + j, pending);
        unsigned long long spins = 0;
        int v;
        while ((v = ld_acquire(prog + j - 1)) < need)
          if (++spins > CH_SPIN_LIMIT) { s_ok = 0; atomicExch(a.err, 1); break; }
        s_seen = v;
      }
      pending = 0;
      __syncthreads();
      seen = s_seen;
      return s_ok != 0;
    };
    bool alive = wait_prev(2);
    if (!alive) { if (tid == 0) st_release(prog + j, CH_DONE); break; }
    // first reflector of the sweep: column j, rows lo..hi (every warp computes it: lane = row)
    double v, tau, beta;
    {
      const double x = (lane < len) ? __ldcg(bw + (size_t)(lo + lane) * CLD + (lane + 1)) : 0.0;
      warp_reflector(x, lane, len, v, tau, beta);
      if (part == 0 && lane < len) __stcg(bw + (size_t)(lo + lane) * CLD + (lane + 1), (lane == 0) ? beta : 0.0);
    }
    while (true) {
      const int lo2 = hi + 1, hi2 = min(hi + CB, n - 1);
      const int len2 = hi2 - lo2 + 1;
      const int off = lo2 - lo;
      const bool has_s = len2 > 0, has_refl = len2 >= 2;
      // ---------------- loads: warp `part` fetches rows c0 .. c0 + CC_W - 1 of both blocks
      double dr[CC_W], sr[CC_W];
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        dr[i] = (l < len && lane <= l) ? __ldcg(bw + (size_t)(lo + l) * CLD + lane) : 0.0;
        sr[i] = (l < len2 && lane < len) ? __ldcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane)) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        if (lane <= l) {
          D[l * CS_ + (l - lane)] = dr[i];
          D[(l - lane) * CS_ + l] = dr[i];
        }
        S[l * CS_ + lane] = sr[i];
      }
      if (part == 0) vs[lane] = v;
      __syncthreads();                                                       // #1
      if (pending && tid == 0) st_release(prog + j, pending);                // hops up to `pending` are complete
      pending = 0;
      // ---------------- partial products over this thread's columns: D v and S v
      {
        double pp = 0.0, uu = 0.0;
#pragma unroll
        for (int c = 0; c < CC_W; ++c) {
          const double vc = vs[c0 + c];
          pp = fma(D[lane * CS_ + c0 + c], vc, pp);
          uu = fma(S[lane * CS_ + c0 + c], vc, uu);
        }
        red[part * CB + lane] = pp;
        red2[part * CB + lane] = uu;
      }
      __syncthreads();                                                       // #2
      double p = 0.0, u = 0.0;
#pragma unroll
      for (int q = 0; q < CC_PARTS; ++q) { p += red[q * CB + lane]; u += red2[q * CB + lane]; }
      p *= tau;
      u *= tau;
      const double al = -0.5 * tau * warp_sum(p * v);
      const double wv = p + al * v;
      if (part == 0) ws[lane] = wv;
      __syncthreads();                                                       // #3
      // ---------------- D <- H D H and S <- S H on this thread's row slice; column 0 of S is replaced below, so it is
      // left alone here (every warp still reads its pre-update value)
      double v2 = 0.0, tau2 = 0.0;
      {
        const double x0 = S[lane * CS_] - u * vs[0];       // column 0 after S H
#pragma unroll
        for (int c = 0; c < CC_W; ++c) {
          const double vc = vs[c0 + c];
          D[lane * CS_ + c0 + c] -= v * ws[c0 + c] + wv * vc;
          if (c0 + c != 0) S[lane * CS_ + c0 + c] -= u * vc;
        }
        double nc0 = x0;
        if (has_refl) {
          double beta2;
          warp_reflector(x0, lane, len2, v2, tau2, beta2);   // identical in every warp
          nc0 = (lane == 0) ? beta2 : 0.0;
        }
        if (part == 0) { v2s[lane] = v2; col0[lane] = nc0; }
      }
      __syncthreads();                                                       // #4
      // ---------------- store D (rows c0 .. of this warp, coalesced) ; partial products v2^T S of this thread's rows
#pragma unroll
      for (int i = 0; i < CC_W; ++i) {
        const int l = c0 + i;
        if (l < len && lane <= l) __stcg(bw + (size_t)(lo + l) * CLD + lane, D[l * CS_ + (l - lane)]);
      }
      double sv[CC_W];
      {
        double zz = 0.0;
#pragma unroll
        for (int i = 0; i < CC_W; ++i) {
          sv[i] = S[(c0 + i) * CS_ + lane];
          zz = fma(v2s[c0 + i], sv[i], zz);
        }
        red[part * CB + lane] = zz;
      }
      __syncthreads();                                                       // #5
      if (has_s) {
        double z = 0.0;
#pragma unroll
        for (int q = 0; q < CC_PARTS; ++q) z += red[q * CB + lane];
        z = (has_refl && lane >= 1) ? z * tau2 : 0.0;
#pragma unroll
        for (int i = 0; i < CC_W; ++i) {
          const int l = c0 + i;
          const double val = (lane == 0) ? col0[l] : sv[i] - v2s[l] * z;
          if (l < len2 && lane < len) __stcg(bw + (size_t)(lo2 + l) * CLD + (off + l - lane), val);
        }
      }
      ++h;
      if (!has_refl) {
        // last hop of the sweep: publish DONE once everybody's stores are issued
        __syncthreads();
        if (tid == 0) st_release(prog + j, CH_DONE);
        break;
      }
      pending = h;                     // published after barrier #1 of the next hop (or by wait_prev)
      v = v2; tau = tau2; lo = lo2; hi = hi2; len = len2;
      alive = wait_prev(h + 2);
      if (!alive) { if (tid == 0) st_release(prog + j, CH_DONE); break; }
    }
    if (!alive) break;
  }
}

__global__ void chase_extract_kernel(const double* __restrict__ bandW, int n, const int* __restrict__ err,
                                     double* __restrict__ d, double* __restrict__ e) {
  const int t = blockIdx.y;
  const double* bw = bandW + (size_t)t * n * CLD;
  // a timed-out sweep leaves the band half reduced: poison the result so that every consumer reports NaN
  const double poison = (err && *err) ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    d[(size_t)t * n + i] = bw[(size_t)i * CLD] + poison;
    e[(size_t)t * n + i] = ((i + 1 < n) ? bw[(size_t)(i + 1) * CLD + 1] : 0.0) + poison;
  }
}

// progress counters (T x n) + ticket counters (T) + error flag, zeroed before every launch
size_t chase_prog_ints(const Plan* pl, int T) { return (size_t)T * pl->n + (size_t)T + 64; }
size_t chase_workspace(const Plan* pl, int T) { return align_up(chase_prog_ints(pl, T) * sizeof(int), 256) + 256; }

int chase_device_init(Plan* pl) {
  // 2: CTA per sweep, 5 barriers per hop (default); 1: CTA per sweep, first formulation; 0: one warp per sweep
  const int variant = getenv("CHD_CHASE") ? atoi(getenv("CHD_CHASE")) : 2;
  pl->chase_variant = variant;
  const size_t smem = variant == 0 ? (size_t)CH_WARPS * (2 * CB * CS_ + 3 * CB) * sizeof(double)
                    : variant == 2 ? (size_t)C2_SMEM_DOUBLES * sizeof(double)
                                   : (size_t)CC_SMEM_DOUBLES * sizeof(double);
  const int threads = variant == 0 ? CH_WARPS * 32 : CC_THREADS;
  int nb = 0;
  if (variant == 0) {
    CHD_CUDA(cudaFuncSetAttribute(chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chase_kernel, threads, smem));
  } else if (variant == 2) {
    CHD_CUDA(cudaFuncSetAttribute(chase_cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chase_cta2_kernel, threads, smem));
  } else {
    CHD_CUDA(cudaFuncSetAttribute(chase_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chase_cta_kernel, threads, smem));
  }
  pl->chase_ctas_per_sm = nb > 0 ? nb : 1;
  // L2 budget for the work bands of one launch (default: everything in one launch; CHD_CHASE_L2_MB to cap) and an
  // explicit cap on matrices per launch (CHD_CHASE_WAVE)
  pl->chase_l2_budget = (size_t)1 << 40;
  if (const char* e = getenv("CHD_CHASE_L2_MB")) pl->chase_l2_budget = (size_t)atoi(e) << 20;
  pl->chase_wave = getenv("CHD_CHASE_WAVE") ? atoi(getenv("CHD_CHASE_WAVE")) : 0;
  if (const char* e = getenv("CHD_CHASE_CTAS_PER_SM")) {   // tuning: leave room for another stream's kernels
    const int v = atoi(e);
    if (v >= 1 && v < pl->chase_ctas_per_sm) pl->chase_ctas_per_sm = v;
  }
  return CHD_OK;
}

int chase_launch(const Plan* pl, double* bandW, int* prog, double* d, double* e, int T, cudaStream_t st) {
  const int n = pl->n;
  const int variant = pl->chase_variant;
  const size_t smem = variant == 0 ? (size_t)CH_WARPS * (2 * CB * CS_ + 3 * CB) * sizeof(double)
                    : variant == 2 ? (size_t)C2_SMEM_DOUBLES * sizeof(double)
                                   : (size_t)CC_SMEM_DOUBLES * sizeof(double);
  const int threads = variant == 0 ? CH_WARPS * 32 : CC_THREADS;
  const int sweeps_per_cta = variant == 0 ? CH_WARPS : 1;
  int* ticket = prog + (size_t)T * n;
  int* err = ticket + T;
  if (n >= 3) {
    CHD_CUDA(cudaMemsetAsync(prog, 0, chase_prog_ints(pl, T) * sizeof(int), st));
    // sweeps in flight per matrix: what fits on the device for all T matrices at once, at most the pipeline depth
    // n / (2 NB).  One launch: sweeps are drawn from the ticket counters, so CTAs that are not resident yet are
    // never waited for.
    // matrices per launch: all of them, unless their work bands (n x 66 doubles each) would overflow L2 -- then the
    // band lines thrash to DRAM between consecutive sweeps (ncu: 184 GB written for 32 matrices at n = 8192) --
    // in which case they are chased in waves that fit
    int wave = T;
    {
      const size_t band_bytes = (size_t)n * CLD * sizeof(double);
      const size_t fit = pl->chase_l2_budget / (band_bytes ? band_bytes : 1);
      if (fit >= 1 && (size_t)wave > fit) wave = (int)fit;
      if (pl->chase_wave > 0 && pl->chase_wave < wave) wave = pl->chase_wave;
    }
    const int cap = pl->sm_count * pl->chase_ctas_per_sm;        // CTAs
    const int want = (n / (2 * CB) + sweeps_per_cta) / sweeps_per_cta;
    int prc;
    if ((prc = profile_mark_begin(PROF_CHASE, st))) return prc;
    for (int t0 = 0; t0 < T; t0 += wave) {
      const int cnt = (T - t0 < wave) ? T - t0 : wave;
      int per = cap / cnt;
      if (per > want) per = want;
      if (per < 1) per = 1;
      ChaseArgs a;
      a.bandW = bandW + (size_t)t0 * n * CLD; a.prog = prog + (size_t)t0 * n; a.ticket = ticket + t0; a.err = err;
      a.n = n;
      if (variant == 0) chase_kernel<<<dim3(per, cnt), threads, smem, st>>>(a);
      else if (variant == 2) chase_cta2_kernel<<<dim3(per, cnt), threads, smem, st>>>(a);
      else chase_cta_kernel<<<dim3(per, cnt), threads, smem, st>>>(a);
      CHD_LAUNCHED();
    }
    if ((prc = profile_mark_end(PROF_CHASE, st))) return prc;
  }
  chase_extract_kernel<<<dim3((n + 255) / 256, T), 256, 0, st>>>(bandW, n, n >= 3 ? err : nullptr, d, e);
  CHD_LAUNCHED();
  return CHD_OK;
}

}  // namespace chd
