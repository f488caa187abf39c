// SPGG lattice step for sm_100a - lattice-resident cluster kernel (small lattices).
//
// The reference's own workloads are small lattices run for very many iterations
// (default_config.yaml: L=100, 100001 iterations; figure sweeps: L=100/200).
// At that size the two-kernels-per-iteration scheme of spgg_kernels.cuh / spgg_fast.cuh is
// bound by launch latency, not by memory.  Here ONE thread-block cluster owns one replica
// for a whole chunk of iterations:
//
//   * the lattice state - Q (16 B/site, four float4 planes), reputation, reward codes and
//     strategies (byte planes, ping-pong) - lives in the shared memory of the cluster's CTAs
//     (row blocks), loaded from HBM once per chunk and written back once;
//   * the two ghost rows a CTA needs from each neighbour are PUSHED into the neighbour's
//     shared memory through DSMEM by the thread that produces them;
//   * the lattice-global max |reward difference| (spgg.py:488) is an all-to-all of one float
//     per CTA through DSMEM; the early-exit test (spgg.py:405) the same with one counter;
//   * two cluster barriers per iteration replace two kernel launches;
//   * statistics: one transposing butterfly per warp for all 28 values, block rows folded by rank
//     0 inside the next barrier window into an on-chip ring of raw rows (k_resident_rows turns
//     them into the public layout once per chunk).
//
// GRID mode (template flag) spreads ONE lattice of up to 1036 x 1036 sites over a cooperative grid
// of one CTA per SM with the same quad code: ghost rows, maxima and counters travel through L2.
//
// Arithmetic, Philox counters and statistics are those of k_step<ModeF32I8> / k_gmax
// (Q-learning, algorithms.py:96-133), so S, R and Q are bit-identical to the other two paths.
#pragma once
#include <cooperative_groups.h>

#include <type_traits>

#include "spgg_kernels.cuh"

namespace spgg {
namespace cg = cooperative_groups;

// debug builds only (-DSPGG_RES_TRACE): cycle stamps of one thread at the phase boundaries of every
// iteration, accumulated per phase into RArgs::trace[rank][8]
#ifdef SPGG_RES_TRACE
#define RES_STAMP(slot)                                                     \
  do {                                                                      \
    if (tid == 0) {                                                         \
      const long long now_ = clock64();                                     \
      trace_acc[slot] += now_ - trace_t;                                    \
      trace_t = now_;                                                       \
    }                                                                       \
  } while (0)
#else
#define RES_STAMP(slot) do {} while (0)
#endif

constexpr int RES_THREADS = 512;  // upper bound; small lattices launch fewer
constexpr int RES_CS_MAX = 16;    // largest cluster (8 is the portable size, 16 needs the non-portable opt-in)
constexpr int RG = 4;             // ghost columns (bytes) on each side of a shared-memory plane row
constexpr int RES_NI = 18;        // integer statistics per thread
constexpr int RES_NF = 10;        // fp32 statistics per thread
constexpr int RES_NRED = RES_NI + RES_NF;
constexpr int RES_RAW = RES_NRED + 1;  // raw statistics row: the 28 lattice sums + the global max
constexpr int RES_RING = 16;           // raw rows kept on chip between flushes to HBM
constexpr int RES_RAW_GMAX = 39;       // column of a raw row in HBM that carries the global max

struct ResGeom {
  int CS;        // CTAs per cluster = row blocks per replica
  int rows_max;  // rows of the largest block: ceil(L / CS)
  int prow;      // plane rows: rows_max + 4 (two ghost rows each side)
  int pitch;     // bytes per plane row: RG + roundup(L,4) + RG, rounded up to 16
  int QR;        // quads (4 consecutive sites) per row: ceil(L / 4)
  int threads;
  int grid;      // 1: grid mode (no reward plane, no cluster mailboxes in shared memory)
};

// byte offsets of the per-CTA shared-memory planes (identical in every CTA of a cluster, so a
// local offset is valid in a neighbour's window)
struct ResSmem {
  size_t q, val, code[2], R[2], C[2], N, tab, redf, redi, gmx, nsel, part, ring, total;
  __host__ __device__ explicit ResSmem(const ResGeom &rg) {
    const size_t nq = (size_t)rg.rows_max * rg.QR * 16;  // one float4 plane (one Q entry, 4 sites per element)
    const size_t nb = (size_t)rg.prow * rg.pitch;
    size_t o = 0;
    q = o; o += 4 * nq;
    val = o; o += rg.grid ? 0 : 4 * nb;   // grid mode looks rewards up from the codes instead (room for 1000 x 1000)
    for (int i = 0; i < 2; ++i) { code[i] = o; o += nb; }
    for (int i = 0; i < 2; ++i) { R[i] = o; o += nb; }
    for (int i = 0; i < 2; ++i) { C[i] = o; o += nb; }
    N = o; o += nb;
    tab = o; o += 256 * sizeof(float);
    redf = o; o += sizeof(float) * (RES_THREADS / 32) * 32;
    redi = o; o += sizeof(unsigned) * (RES_NI + 2);
    gmx = o; o += rg.grid ? 0 : sizeof(float) * RES_CS_MAX;
    nsel = o; o += rg.grid ? 0 : sizeof(unsigned) * RES_CS_MAX;
    o = (o + 15) / 16 * 16;
    part = o; o += rg.grid ? 0 : 2 * sizeof(double) * RES_CS_MAX * RES_NRED;  // two buffers, by iteration parity
    ring = o; o += rg.grid ? 0 : sizeof(double) * RES_RING * RES_RAW;         // finished raw rows waiting for their flush (rank 0)
    total = (o + 15) / 16 * 16;
  }
};

struct RArgs {
  Geom g;
  ResGeom rg;
  const RepConst *rc;
  void *Q;            // float4[n_rep][L*L]
  void *R;            // int8 planes holding the current state (read at start, written at end)
  uint32_t *S;        // strategy bit planes, same
  double *stats;      // [n_rep][cap][NSTAT]
  int *stop_at;       // [n_rep]
  const uint32_t *thr_tab;  // [cap+1][n_rep]: ceil(eps*2^24) used at iteration t0 + idx
  int t0;             // iterations completed before this launch
  int n_steps;
  int cap;
  // grid mode (one lattice over every SM, cooperative launch): global images of the byte planes the
  // neighbours write ghost rows into, per-block statistics rows, per-iteration counters / maxima
  unsigned char *gimg;   // [CS][6 * prow * pitch]
  double *gpart;         // [RES_RING][CS][RES_NRED]
  unsigned *gnsel;       // [cap]   cooperating actions chosen in iteration t0 + idx
  float *gmaxtab;        // [n_rep][cap] lattice-global maxima (zeroed per chunk)
#ifdef SPGG_RES_TRACE
  long long *trace;   // [n_rep * CS][8] accumulated cycles per phase
#endif
};

// Rewards around a quad (4 consecutive sites of a row, first plane index `base`) held in
// registers: c[i] = own row, column base-2+i; u[i] / d[i] = row above / below, column base-1+i;
// u2[i] / d2[i] = two rows above / below, column base+i.  M=1 touches c[1..6], u[1..4], d[1..4].
template <int M>
struct ValWin {
  float c[8], u[6], d[6], u2[4], d2[4];
};
template <int M, bool BELOW>
__device__ __forceinline__ void load_win(ValWin<M> &w, const float *V, int base, int pitch) {
  const float4 c4 = *reinterpret_cast<const float4 *>(V + base);
  w.c[2] = c4.x; w.c[3] = c4.y; w.c[4] = c4.z; w.c[5] = c4.w;
  w.c[1] = V[base - 1];
  const float4 u4 = *reinterpret_cast<const float4 *>(V + base - pitch);
  w.u[1] = u4.x; w.u[2] = u4.y; w.u[3] = u4.z; w.u[4] = u4.w;
  if constexpr (M == 2) {
    w.c[0] = V[base - 2];
    w.u[0] = V[base - pitch - 1];
    w.u[5] = V[base - pitch + 4];
    const float4 t4 = *reinterpret_cast<const float4 *>(V + base - 2 * pitch);
    w.u2[0] = t4.x; w.u2[1] = t4.y; w.u2[2] = t4.z; w.u2[3] = t4.w;
  }
  if constexpr (BELOW) {
    w.c[6] = V[base + 4];
    const float4 d4 = *reinterpret_cast<const float4 *>(V + base + pitch);
    w.d[1] = d4.x; w.d[2] = d4.y; w.d[3] = d4.z; w.d[4] = d4.w;
    if constexpr (M == 2) {
      w.c[7] = V[base + 5];
      w.d[0] = V[base + pitch - 1];
      w.d[5] = V[base + pitch + 4];
      const float4 t4 = *reinterpret_cast<const float4 *>(V + base + 2 * pitch);
      w.d2[0] = t4.x; w.d2[1] = t4.y; w.d2[2] = t4.z; w.d2[3] = t4.w;
    }
  }
}
// the same window looked up from the reward codes (grid mode keeps no reward plane): byte b of the
// code plane -> tab[b >> 1]
template <int M, bool BELOW>
__device__ __forceinline__ void load_win_code(ValWin<M> &w, const uint8_t *code, const float *tab, int base,
                                              int pitch) {
  auto row4 = [&](int at, float *dst) {
    const uint32_t c = *reinterpret_cast<const uint32_t *>(code + at);
    dst[0] = tab[(c >> 1) & 127u]; dst[1] = tab[(c >> 9) & 127u];
    dst[2] = tab[(c >> 17) & 127u]; dst[3] = tab[c >> 25];
  };
  auto one = [&](int at) { return tab[code[at] >> 1]; };
  row4(base, &w.c[2]);
  w.c[1] = one(base - 1);
  row4(base - pitch, &w.u[1]);
  if constexpr (M == 2) {
    w.c[0] = one(base - 2);
    w.u[0] = one(base - pitch - 1);
    w.u[5] = one(base - pitch + 4);
    row4(base - 2 * pitch, &w.u2[0]);
  }
  if constexpr (BELOW) {
    w.c[6] = one(base + 4);
    row4(base + pitch, &w.d[1]);
    if constexpr (M == 2) {
      w.c[7] = one(base + 5);
      w.d[0] = one(base + pitch - 1);
      w.d[5] = one(base + pitch + 4);
      row4(base + 2 * pitch, &w.d2[0]);
    }
  }
}
// reward of neighbour z (order of c_off: the value at (row - dx, col - dy)) of site k of the quad
template <int M>
__device__ __forceinline__ float win_nbr(const ValWin<M> &w, int k, int z) {
  switch (z) {
    case 0: return w.u[1 + k];
    case 1: return w.d[1 + k];
    case 2: return w.c[1 + k];
    case 3: return w.c[3 + k];
    case 4: return w.u2[k];
    case 5: return w.d2[k];
    case 6: return w.c[k];
    case 7: return w.c[4 + k];
    case 8: return w.u[k];
    case 9: return w.u[2 + k];
    case 10: return w.d[k];
    default: return w.d[2 + k];
  }
}

// FULL: L is a multiple of 4, so every quad is complete (no per-site validity branches: the four
// sites of a quad are independent instruction streams the scheduler can interleave) and the
// periodic column images are whole words.
// GRID: instead of one cluster per replica, ONE lattice is spread over a cooperative grid of up to
// one CTA per SM (lattices up to about 1100 x 1100): ghost rows travel through L2 (each block owns
// a global image of its byte planes; neighbours write its ghost rows there and it pulls them after
// the barrier), the maximum and the early-exit counter are global atomics, the barriers grid-wide.
template <int M, bool ACTION, bool FULL, bool GRID = false>
__global__ void __launch_bounds__(RES_THREADS) k_resident(RArgs a) {
  constexpr int NK = (M == 2) ? 12 : 4;
  cg::cluster_group cluster = cg::this_cluster();
  const Geom &g = a.g;
  const ResGeom &rg = a.rg;
  const int CS = rg.CS;
  const int rep = blockIdx.x / CS;
  const int rank = GRID ? (int)(blockIdx.x % CS) : (int)cluster.block_rank();
  cg::grid_group grid = cg::this_grid();
  const int L = g.L, pitch = rg.pitch, QR = rg.QR;
  const int W = pitch >> 2;  // 32-bit words per plane row
  const int base_rows = L / CS, rem = L % CS;
  const int nrow = base_rows + (rank < rem ? 1 : 0);
  const int row_start = rank * base_rows + min(rank, rem);
  const int up = (rank + CS - 1) % CS, down = (rank + 1) % CS;
  const int nrow_up = base_rows + (up < rem ? 1 : 0);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

  extern __shared__ __align__(16) unsigned char smem[];
  const ResSmem lay(rg);
  float4 *sQ = reinterpret_cast<float4 *>(smem + lay.q);
  const int q_plane = rg.rows_max * QR;  // float4 elements per Q plane
  float *s_val = reinterpret_cast<float *>(smem + lay.val);
  uint8_t *s_N = smem + lay.N;
  float *s_tab = reinterpret_cast<float *>(smem + lay.tab);
  float *s_ratio = s_tab + 128;
  float *s_redf = reinterpret_cast<float *>(smem + lay.redf);
  unsigned *s_redi = reinterpret_cast<unsigned *>(smem + lay.redi);  // [RES_NI] = block max of the reward differences
  float *s_gmx = reinterpret_cast<float *>(smem + lay.gmx);
  unsigned *s_nsel = reinterpret_cast<unsigned *>(smem + lay.nsel);
  double *s_part = reinterpret_cast<double *>(smem + lay.part);
  const int o_code0 = (int)lay.code[0], o_R0 = (int)lay.R[0], o_C0 = (int)lay.C[0];
  const int nb = rg.prow * pitch;  // bytes per byte plane; set 1 follows set 0
  __shared__ RepConst s_rc;
  __shared__ int s_noff[12];  // plane-index offset of neighbour z: dx*pitch + dy
  if (threadIdx.x < 12) s_noff[threadIdx.x] = c_off[threadIdx.x][0] * pitch + c_off[threadIdx.x][1];

  // ---- zero every plane (cells that are neither sites nor ghosts stay zero for good: the
  // byte-parallel group count below must not see junk), constants, accumulators
  for (size_t i = (size_t)tid * 16; i < lay.total; i += (size_t)nthr * 16)
    *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (int)(sizeof(RepConst) / 4); i += nthr)
    reinterpret_cast<uint32_t *>(&s_rc)[i] = reinterpret_cast<const uint32_t *>(a.rc + rep)[i];
  __syncthreads();
  for (int i = tid; i < 128; i += nthr) {
    s_tab[i] = s_rc.rewtab[i];
    s_ratio[i] = s_rc.ratiotab[i];
  }
  const RepConst &rc = s_rc;

  // ---- load the state of this row block (+ two ghost rows each side, two ghost columns)
  float4 *Qg = reinterpret_cast<float4 *>(a.Q) + (long long)rep * g.site_stride;
  int8_t *Rg = reinterpret_cast<int8_t *>(a.R) + (long long)rep * g.plane_stride;
  uint32_t *Sg = a.S + (long long)rep * g.bits_stride;
  {
    uint8_t *R0 = smem + o_R0, *C0 = smem + o_C0;
    const int wc = L + 4, ncell = (nrow + 4) * wc;
    for (int e = tid; e < ncell; e += nthr) {
      const int rr = e / wc - 2, cc = e % wc - 2;
      int grow = row_start + rr, gcol = cc;
      grow += (grow < 0) ? L : 0; grow -= (grow >= L) ? L : 0;
      gcol += (gcol < 0) ? L : 0; gcol -= (gcol >= L) ? L : 0;
      const int idx = (rr + 2) * pitch + RG + cc;
      R0[idx] = (uint8_t)Rg[(long long)(grow + GH) * g.pitchB + CPAD + gcol];
      const uint32_t w = Sg[(long long)(grow + GH) * g.pitchW + WPAD + (gcol >> 5)];
      C0[idx] = (uint8_t)(((w >> (gcol & 31)) & 1u) ^ 1u);
    }
    float *sQf = reinterpret_cast<float *>(sQ);
    for (int e = tid; e < nrow * L; e += nthr) {
      const int rr = e / L, cc = e % L;
      const float4 v = Qg[(long long)(row_start + rr) * L + cc];
      const int o = (rr * QR + (cc >> 2)) * 4 + (cc & 3);
      sQf[o] = v.x;
      sQf[(q_plane * 4) + o] = v.y;
      sQf[(q_plane * 8) + o] = v.z;
      sQf[(q_plane * 12) + o] = v.w;
    }
  }
  // where the ghost rows this block produces for its neighbours go: the neighbour's shared memory
  // (DSMEM), or - grid mode - the neighbour's global plane image (same offsets as in shared memory)
  const size_t img_stride = (size_t)6 * nb;
  unsigned char *smem_up, *smem_dn;
  if constexpr (GRID) {
    smem_up = a.gimg + (size_t)up * img_stride - o_code0;
    smem_dn = a.gimg + (size_t)down * img_stride - o_code0;
  } else {
    smem_up = cluster.map_shared_rank(smem, up);
    smem_dn = cluster.map_shared_rank(smem, down);
  }
  __syncthreads();
  if constexpr (!GRID) cluster.sync();  // every CTA of the cluster runs and has zeroed its planes before anyone pushes ghosts

  // ---- loop-invariant iteration patterns: plane words (row, word) and quads (row, quad)
  const int w_r0 = tid / W, w_c0 = tid % W, w_dr = nthr / W, w_dc = nthr % W;
  const int q_r0 = tid / QR, q_c0 = tid % QR, q_dr = nthr / QR, q_dc = nthr % QR;
  const long long n_sites = (long long)L * L;

  // Raw statistics row of iteration fs (rank 0, warp 0): the 28 lattice sums (block rows folded in
  // a fixed order) and the global max.  It is not on the critical path - the row of iteration s is
  // folded inside the window of the NEXT iteration's max-exchange barrier (arrive - fold - wait),
  // while the other blocks' arrivals are in flight - and it does not touch HBM: rows collect in a
  // shared-memory ring that is flushed every RES_RING iterations (a cluster barrier waits for the
  // thread's outstanding global stores).  k_resident_rows turns raw rows into the public layout.
  double *s_ring = reinterpret_cast<double *>(smem + lay.ring);
  int ring_first = -1, ring_last = -1;  // iterations of the rows waiting in the ring
  auto flush_ring = [&]() {
    if (ring_first < 0) return;
    for (int r = ring_first; r <= ring_last; ++r) {
      double *row = a.stats + ((long long)rep * a.cap + r) * NSTAT;
      if (lane < RES_NRED) row[lane] = s_ring[(r % RES_RING) * RES_RAW + lane];
      if (lane == RES_NRED) row[RES_RAW_GMAX] = s_ring[(r % RES_RING) * RES_RAW + RES_NRED];
    }
    ring_first = -1;
  };
  auto do_fold = [&](int fs, float fgm) {
    double f = 0.0;
    if (lane < RES_NRED)
      for (int k = 0; k < CS; ++k) f += s_part[(fs & 1) * (RES_CS_MAX * RES_NRED) + k * RES_NRED + lane];
    if (lane == RES_NRED) f = (double)fgm;
    if (lane <= RES_NRED) s_ring[(fs % RES_RING) * RES_RAW + lane] = f;
    if (ring_first < 0) ring_first = fs;
    ring_last = fs;
    __syncwarp();
    if (fs - ring_first == RES_RING - 1) flush_ring();
  };
  // grid mode: the per-block rows of up to RES_RING iterations wait in HBM; block r folds the r-th
  // pending iteration (fixed order over the blocks: deterministic) straight into its raw row
  auto grid_fold = [&](int first, int last) {
    const int fs = first + rank;
    if (fs > last) return;  // block-uniform
    // up to 16 warps read interleaved slices of the block rows (independent L2 loads in flight),
    // warp 0 adds the slice sums in order; the N plane is free between two iterations: scratch
    double *scr = reinterpret_cast<double *>(s_N);
    const int P = min(nwarp, 16);
    if (warp < P && lane < RES_NRED) {
      const double *pp = a.gpart + ((size_t)(fs % RES_RING) * CS) * RES_NRED + lane;
      double f = 0.0;
#pragma unroll 4
      for (int k = warp; k < CS; k += P) f += __ldcg(pp + (size_t)k * RES_NRED);
      scr[warp * 32 + lane] = f;
    }
    __syncthreads();
    if (warp == 0) {
      double f = 0.0;
      if (lane < RES_NRED)
        for (int j = 0; j < P; ++j) f += scr[j * 32 + lane];
      double *row = a.stats + ((long long)rep * a.cap + fs) * NSTAT;
      if (lane < RES_NRED) row[lane] = f;
      if (lane == RES_NRED) row[RES_RAW_GMAX] = (double)__ldcg(a.gmaxtab + (long long)rep * a.cap + fs);
    }
    __syncthreads();  // the scratch is the N plane again
  };
  int gfold_first = 0;  // grid mode: first iteration whose block rows have not been folded yet
  int pend_s = -1, last_s = -1;
  float pend_gm = 0.0f;

  int stop = a.stop_at[rep];
  int cur = 0;
#ifdef SPGG_RES_TRACE
  long long trace_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long trace_t = clock64();
#endif
  for (int s = 0; s <= a.n_steps; ++s) {
    const int j = a.t0 + s;
    if (stop >= 0 && j > stop) break;
    const bool upd = s > 0;
    const bool sel = (s < a.n_steps) && !(stop >= 0 && j == stop);
    const uint32_t thr = sel ? __ldg(a.thr_tab + (long long)(s + 1) * g.n_rep + rep) : 0u;
    const uint8_t *codeC = smem + o_code0 + cur * nb;
    const uint8_t *Rc = smem + o_R0 + cur * nb;
    const uint8_t *Cc = smem + o_C0 + cur * nb;
    const int o_code_n = o_code0 + (cur ^ 1) * nb, o_R_n = o_R0 + (cur ^ 1) * nb, o_C_n = o_C0 + (cur ^ 1) * nb;

    // ---- phase 1: cooperators per group (spgg.py:23-36) for the block and a one-site ring,
    // four sites per word; reward of every site the block can see (own rows + M ghost rows)
    {
      const uint32_t *Cw = reinterpret_cast<const uint32_t *>(Cc);
      uint32_t *Nw = reinterpret_cast<uint32_t *>(s_N);
      int rr = w_r0, wq = w_c0;
      while (rr < nrow + 2) {  // plane rows 1 .. nrow+2  (local rows -1 .. nrow)
        const int i4 = (rr + 1) * W + wq;
        const uint32_t c = Cw[i4];
        const uint32_t lw = __funnelshift_l(Cw[i4 - 1], c, 8), rw = __funnelshift_r(c, Cw[i4 + 1], 8);
        Nw[i4] = c + Cw[i4 - W] + Cw[i4 + W] + lw + rw;
        rr += w_dr; wq += w_dc;
        if (wq >= W) { wq -= W; rr += 1; }
      }
      if (upd && !GRID) {
        const uint32_t *cw = reinterpret_cast<const uint32_t *>(codeC);
        float4 *vw = reinterpret_cast<float4 *>(s_val);
        rr = w_r0; wq = w_c0;
        while (rr < nrow + 2 * M) {  // plane rows 2-M .. nrow+1+M
          const int i4 = (rr + 2 - M) * W + wq;
          const uint32_t c = cw[i4];
          vw[i4] = make_float4(s_tab[(c >> 1) & 127u], s_tab[(c >> 9) & 127u], s_tab[(c >> 17) & 127u],
                               s_tab[c >> 25]);
          rr += w_dr; wq += w_dc;
          if (wq >= W) { wq -= W; rr += 1; }
        }
      }
    }
    __syncthreads();
    RES_STAMP(0);  // phase 1 + barrier

    // ---- phase 2: lattice-global max |reward difference| (spgg.py:486-488): each unordered
    // neighbour pair once, block max, all-to-all through DSMEM
    float inv_den = 0.0f, gm = 0.0f;
    if (upd) {
      float lmax = 0.0f;
      int rr = q_r0, qc = q_c0;
      while (rr < nrow) {
        const int base = (rr + 2) * pitch + RG + 4 * qc;
        ValWin<M> w;
        if constexpr (GRID) load_win_code<M, false>(w, codeC, s_tab, base, pitch);
        else load_win<M, false>(w, s_val, base, pitch);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (FULL || 4 * qc + k < L) {
            const float vx = w.c[2 + k];
#pragma unroll
            for (int z = 0; z < NK; ++z) {
              if (z == 1 || z == 3 || z == 5 || z == 7 || z == 10 || z == 11) continue;  // mirror images
              const float d = fabsf(__fsub_rn(win_nbr<M>(w, k, z), vx));
              lmax = d > lmax ? d : lmax;
            }
          }
        }
        rr += q_dr; qc += q_dc;
        if (qc >= QR) { qc -= QR; rr += 1; }
      }
      const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(lmax));
      if (lane == 0) atomicMax(&s_redi[RES_NI], wm);
      __syncthreads();
      RES_STAMP(1);  // phase 2 + block max
      if constexpr (GRID) {
        if (tid == 0)  // non-negative IEEE floats order like unsigned integers; the table is zeroed per chunk
          atomicMax(reinterpret_cast<unsigned *>(a.gmaxtab) + (long long)rep * a.cap + s, s_redi[RES_NI]);
        grid.sync();
        RES_STAMP(2);
        gm = __ldcg(a.gmaxtab + (long long)rep * a.cap + s);
      } else {
        if (tid < CS) {
          float *dst = reinterpret_cast<float *>(cluster.map_shared_rank(smem, tid) + lay.gmx);
          dst[rank] = __uint_as_float(s_redi[RES_NI]);
        }
        cluster.barrier_arrive();
        if (rank == 0 && warp == 0 && pend_s >= 0) do_fold(pend_s, pend_gm);
        pend_s = -1;
        cluster.barrier_wait();
        RES_STAMP(2);  // max exchange + cluster barrier A (+ the previous iteration's statistics row on rank 0)
        for (int k = 0; k < CS; ++k) gm = fmaxf(gm, s_gmx[k]);
      }
      inv_den = __fdiv_rn(1.0f, __fadd_rn(gm, rc.leps_f));  // spgg.py:489 denominator
    }

    // ---- phase 3: finish iteration j (TD + neighbour-aware update, statistics), choose the
    // action of iteration j+1, update the reputation, emit the reward code of j+1
    // packed per-thread counters (a thread visits at most 128 sites per iteration): 8-bit class
    // counts (class = C_old*2 + coop), 16-bit SigmaN sums per class, 10-bit group histogram
    unsigned gnsel_v = 0;
    unsigned pk_n = 0;
    unsigned long long pk_sn = 0, pk_grp = 0;
    unsigned n_best = 0, n_best2 = 0, n_sel_coop = 0;
    int tri = 0;
    float tq[4] = {0.f, 0.f, 0.f, 0.f}, tqc[4] = {0.f, 0.f, 0.f, 0.f};
    float tni = 0.f, tratio = 0.f;
    auto phase3 = [&](auto upd_tag, auto sel_tag) {
      // compile-time copies of the two warp-uniform flags: the steady state (both set) has a
      // branch-free site body
      constexpr bool upd = decltype(upd_tag)::value, sel = decltype(sel_tag)::value;
      int rr = q_r0, qc = q_c0;
      while (rr < nrow) {
        const int base = (rr + 2) * pitch + RG + 4 * qc;
        const int qi = rr * QR + qc;
        float qv[4][4];  // [entry 2s+a][site of the quad]
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          const float4 v = sQ[z * q_plane + qi];
          qv[z][0] = v.x; qv[z][1] = v.y; qv[z][2] = v.z; qv[z][3] = v.w;
        }
        uint32_t w4[4] = {0, 0, 0, 0};
        if constexpr (sel)  // counter = (column / 4, global row, iteration, 0): one call per quad
          philox4x32_10((uint32_t)qc, (uint32_t)(row_start + rr), (uint32_t)(j + 1), 0u, rc.seed_lo,
                        rc.seed_hi, w4);
        // the quad's bytes of every plane as one word each
        const uint32_t Rw = *reinterpret_cast<const uint32_t *>(Rc + base);
        const uint32_t Cq = *reinterpret_cast<const uint32_t *>(Cc + base);
        const uint32_t Nq = *reinterpret_cast<const uint32_t *>(s_N + base);
        uint32_t codeq = 0u;
        if constexpr (upd) codeq = *reinterpret_cast<const uint32_t *>(codeC + base);
        const int nvalid = FULL ? 4 : min(4, L - 4 * qc);
        tri += __dp4a((int)Rw, (int)(0x01010101u >> (8 * (4 - nvalid))), 0);
        // post-action state of iteration j == pre-action state of j+1 (spgg.py:423 vs 409):
        // sign of the reputation summed over the site and its neighbours (spgg.py:292-307),
        // byte-parallel dot products pick the columns each site needs
        int racc[4] = {0, 0, 0, 0};
        if constexpr (!ACTION) {
          constexpr int m1[4] = {0x00000001, 0x00000100, 0x00010000, 0x01000000};
          constexpr int m3[4] = {0x00000101, 0x00010101, 0x01010100, 0x01010000};
          const int Ru = *reinterpret_cast<const int *>(Rc + base - pitch);
          const int Rd = *reinterpret_cast<const int *>(Rc + base + pitch);
          if constexpr (M == 1) {
            racc[0] = (int)(int8_t)Rc[base - 1];
            racc[3] = (int)(int8_t)Rc[base + 4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              racc[k] = __dp4a((int)Rw, m3[k], __dp4a(Ru, m1[k], __dp4a(Rd, m1[k], racc[k])));
          } else {
            constexpr int m5[4] = {0x00010101, 0x01010101, 0x01010101, 0x01010100};
            const int Rl = *reinterpret_cast<const int *>(Rc + base - 4);
            const int Rr = *reinterpret_cast<const int *>(Rc + base + 4);
            const int Rul = *reinterpret_cast<const int *>(Rc + base - pitch - 4);
            const int Rur = *reinterpret_cast<const int *>(Rc + base - pitch + 4);
            const int Rdl = *reinterpret_cast<const int *>(Rc + base + pitch - 4);
            const int Rdr = *reinterpret_cast<const int *>(Rc + base + pitch + 4);
            const int Ru2 = *reinterpret_cast<const int *>(Rc + base - 2 * pitch);
            const int Rd2 = *reinterpret_cast<const int *>(Rc + base + 2 * pitch);
            racc[0] = __dp4a(Rl, 0x01010000, __dp4a(Rul, 0x01000000, __dp4a(Rdl, 0x01000000, 0)));
            racc[1] = __dp4a(Rl, 0x01000000, 0);
            racc[2] = __dp4a(Rr, 0x00000001, 0);
            racc[3] = __dp4a(Rr, 0x00000101, __dp4a(Rur, 0x00000001, __dp4a(Rdr, 0x00000001, 0)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              racc[k] = __dp4a((int)Rw, m5[k], __dp4a(Ru, m3[k], __dp4a(Rd, m3[k],
                        __dp4a(Ru2, m1[k], __dp4a(Rd2, m1[k], racc[k])))));
          }
        }
        ValWin<M> vw;
        if constexpr (upd) {
          if constexpr (GRID) load_win_code<M, true>(vw, codeC, s_tab, base, pitch);
          else load_win<M, true>(vw, s_val, base, pitch);
        }
        uint32_t sw = 0, coopw = 0, rneww = 0;  // new state / action / reputation bytes of the quad
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = 4 * qc + k;
          if (!FULL && col >= L) continue;
          const int idx = base + k;
          float q0_ = qv[0][k], q1_ = qv[1][k], q2_ = qv[2][k], q3_ = qv[3][k];
          const int r_old = (int)(int8_t)((Rw >> (8 * k)) & 0xffu);
          const int Ccur = (int)((Cq >> (8 * k)) & 0xffu);
          const int s_new = ACTION ? Ccur : (racc[k] > 0);
          if constexpr (upd) {
            const unsigned code = (codeq >> (8 * k)) & 0xffu;
            const int sO = code & 1u, coop = (code >> 1) & 1u, wasC = (code >> 2) & 1u;
            const int act = coop ^ 1;
            const float vx = vw.c[2 + k];
            float best = 0.0f;
            int kstar = 0;
#pragma unroll
            for (int z = 0; z < NK; ++z) {  // first arg-max wins, spgg.py:486-494
              const float d = __fsub_rn(win_nbr<M>(vw, k, z), vx);
              if (z == 0 || d > best) { best = d; kstar = z; }
            }
            const bool same = (((codeC[idx - s_noff[kstar]] >> 1) & 1u) == (unsigned)coop);
            const int e = 2 * sO + act;
            const float qe = sel4<float>(e, q0_, q1_, q2_, q3_);
            const float na = s_new ? q2_ : q0_, nb = s_new ? q3_ : q1_;  // pre-update row of s'
            const float td = __fsub_rn(__fmaf_rn(rc.gamma_f, fmaxf(na, nb), vx), qe);   // algorithms.py:125-128
            const float qtd = __fmaf_rn(rc.alpha_f, td, qe);                             // algorithms.py:131
            const float lam = __fmul_rn(__fmul_rn(rc.kappa_f, fmaxf(0.0f, best)), inv_den);  // spgg.py:489
            const float nu = same ? lam : -lam;                                          // spgg.py:494-495
            const float na2 = (s_new == sO && act == 0) ? qtd : na;
            const float nb2 = (s_new == sO && act == 1) ? qtd : nb;
            const float td2 = __fsub_rn(__fmaf_rn(rc.gamma_f, fmaxf(na2, nb2), vx), qtd);
            const float qfin = __fadd_rn(qtd, nu);                                       // spgg.py:509
            const float an = fabsf(nu);
            tni += __fdividef(an, fabsf(rc.alpha_f * td2) + an + 1e-8f) * 100.0f;        // spgg.py:512
            pk_sn += (unsigned long long)(code >> 3) << (16 * (wasC * 2 + coop));
            tratio += s_ratio[code >> 1];  // the table is zero for defecting codes and when w_R = 0
            q0_ = (e == 0) ? qfin : q0_;
            q1_ = (e == 1) ? qfin : q1_;
            q2_ = (e == 2) ? qfin : q2_;
            q3_ = (e == 3) ? qfin : q3_;
            pk_n += 1u << (8 * (wasC * 2 + coop));
            if (best > 0.0f) { n_best += 1u; n_best2 += (kstar >= 4); }
            pk_grp += 1ull << (10 * (5 - (int)((Nq >> (8 * k)) & 0xffu)));               // spgg.py:586-592
            const float m = wasC ? 1.0f : 0.0f;
            tq[0] += q0_; tq[1] += q1_; tq[2] += q2_; tq[3] += q3_;
            tqc[0] = fmaf(m, q0_, tqc[0]); tqc[1] = fmaf(m, q1_, tqc[1]);
            tqc[2] = fmaf(m, q2_, tqc[2]); tqc[3] = fmaf(m, q3_, tqc[3]);
            qv[0][k] = q0_; qv[1][k] = q1_; qv[2][k] = q2_; qv[3][k] = q3_;
          }
          if constexpr (sel) {
            const int explore = (w4[k] >> 8) < thr;                                      // algorithms.py:105
            const int rnd = (int)(w4[k] & 1u);                                           // algorithms.py:108
            const float ga = s_new ? q2_ : q0_, gb = s_new ? q3_ : q1_;
            const int greedy = (gb > ga) ? 1 : 0;                                        // argmax, tie -> C
            const int a_new = explore ? rnd : greedy;
            n_sel_coop += (a_new == 0);
            int t = r_old + (a_new == 0 ? rc.gain_i : -rc.loss_i);                       // spgg.py:321-323
            t = max(t, rc.rmin_i);
            t = min(t, rc.rmax_i);
            sw |= (uint32_t)s_new << (8 * k);
            coopw |= (uint32_t)(a_new ^ 1) << (8 * k);
            rneww |= ((uint32_t)t & 0xffu) << (8 * k);
          }
        }
        if constexpr (sel) {
          // reward code of iteration j+1 for the four sites at once: SigmaN (cooperators summed over
          // the site's five groups, 0..25) << 3 | C_old << 2 | coop << 1 | state
          const uint32_t *Nw32 = reinterpret_cast<const uint32_t *>(s_N + base);
          const uint32_t snw = Nq + Nw32[-W] + Nw32[W] + __funnelshift_l(Nw32[-1], Nq, 8) +
                               __funnelshift_r(Nq, Nw32[1], 8);
          const uint32_t codew = (snw << 3) | (Cq << 2) | (coopw << 1) | sw;
          if constexpr (FULL) {
            // own cells, the periodic column image of the first / last quad of a row, and - for rows
            // within two of a block edge - the neighbour block's ghost rows (DSMEM)
            auto putw = [&](unsigned char *basep, int at) {
              *reinterpret_cast<uint32_t *>(basep + o_code_n + at) = codew;
              *reinterpret_cast<uint32_t *>(basep + o_R_n + at) = rneww;
              *reinterpret_cast<uint32_t *>(basep + o_C_n + at) = coopw;
              if (qc == 0) {
                *reinterpret_cast<uint32_t *>(basep + o_code_n + at + L) = codew;
                *reinterpret_cast<uint32_t *>(basep + o_R_n + at + L) = rneww;
                *reinterpret_cast<uint32_t *>(basep + o_C_n + at + L) = coopw;
              }
              if (qc == QR - 1) {
                *reinterpret_cast<uint32_t *>(basep + o_code_n + at - L) = codew;
                *reinterpret_cast<uint32_t *>(basep + o_R_n + at - L) = rneww;
                *reinterpret_cast<uint32_t *>(basep + o_C_n + at - L) = coopw;
              }
            };
            putw(smem, base);
            if (rr < 2) putw(smem_up, (nrow_up + rr + 2) * pitch + RG + 4 * qc);
            if (rr >= nrow - 2) putw(smem_dn, (rr - nrow + 2) * pitch + RG + 4 * qc);
          } else {
            // L not a multiple of 4: the ghost columns share words with sites, store bytes
            for (int k = 0; k < nvalid; ++k) {
              const int col = 4 * qc + k;
              const uint8_t cnew = (uint8_t)(codew >> (8 * k)), rnew = (uint8_t)(rneww >> (8 * k)),
                            Cnew = (uint8_t)(coopw >> (8 * k));
              const int gdx = (col < GC) ? L : ((col >= L - GC) ? -L : 0);
              auto put = [&](unsigned char *basep, int at) {
                basep[o_code_n + at] = cnew; basep[o_R_n + at] = rnew; basep[o_C_n + at] = Cnew;
                if (gdx) {
                  basep[o_code_n + at + gdx] = cnew; basep[o_R_n + at + gdx] = rnew; basep[o_C_n + at + gdx] = Cnew;
                }
              };
              put(smem, base + k);
              if (rr < 2) put(smem_up, (nrow_up + rr + 2) * pitch + RG + col);
              if (rr >= nrow - 2) put(smem_dn, (rr - nrow + 2) * pitch + RG + col);
            }
          }
        }
        if constexpr (upd) {
#pragma unroll
          for (int z = 0; z < 4; ++z) sQ[z * q_plane + qi] = make_float4(qv[z][0], qv[z][1], qv[z][2], qv[z][3]);
        }
        rr += q_dr; qc += q_dc;
        if (qc >= QR) { qc -= QR; rr += 1; }
      }
    };
    if (upd && sel) phase3(std::true_type{}, std::true_type{});
    else if (upd) phase3(std::true_type{}, std::false_type{});
    else if (sel) phase3(std::false_type{}, std::true_type{});
    else phase3(std::false_type{}, std::false_type{});

    RES_STAMP(3);  // phase 3
    // ---- block statistics.  All 28 per-thread values travel as fp32 (the integer counters of a
    // block stay below 2^24, so their fp32 sums are exact) through ONE transposing butterfly: at
    // each of the five levels a lane keeps half of its values and hands the other half to its
    // partner, so 31 shuffles reduce all values at once and lane z ends up with the warp sum of
    // value z (a reduction per value would cost 28 x 5 dependent shuffles).
    {
      float v[32];
#pragma unroll
      for (int z = 0; z < 4; ++z) {
        v[z] = (float)((pk_n >> (8 * z)) & 0xffu);
        v[4 + z] = (float)(unsigned)((pk_sn >> (16 * z)) & 0xffffull);
        v[18 + z] = tq[z];
        v[22 + z] = tqc[z];
        v[28 + z] = 0.0f;
      }
#pragma unroll
      for (int z = 0; z < 6; ++z) v[8 + z] = (float)(unsigned)((pk_grp >> (10 * z)) & 0x3ffull);
      v[14] = (float)n_best; v[15] = (float)n_best2; v[16] = (float)n_sel_coop; v[17] = (float)tri;
      v[26] = tni; v[27] = tratio;
#pragma unroll
      for (int lvl = 0; lvl < 5; ++lvl) {
        const int o = 16 >> lvl, half = 16 >> lvl;  // partner distance; values kept at this level
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float keep = hi ? v[i + half] : v[i];
          const float send = hi ? v[i] : v[i + half];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      if (lane < RES_NRED) s_redf[warp * 32 + lane] = v[0];
    }
    __syncthreads();
    if (tid < RES_NRED) {
      double x = 0.0;
      for (int w = 0; w < nwarp; ++w) x += (double)s_redf[w * 32 + tid];
      if constexpr (GRID) {
        a.gpart[((size_t)(s % RES_RING) * CS + rank) * RES_NRED + tid] = x;
        if (tid == 16 && x != 0.0) atomicAdd(a.gnsel + s, (unsigned)x);
      } else {
        double *dst = reinterpret_cast<double *>(cluster.map_shared_rank(smem, 0) + lay.part);
        dst[(s & 1) * (RES_CS_MAX * RES_NRED) + rank * RES_NRED + tid] = x;
        if (tid == 16)  // cooperating actions just chosen by this block: to every block (early-exit test)
          for (int k = 0; k < CS; ++k)
            reinterpret_cast<unsigned *>(cluster.map_shared_rank(smem, k) + lay.nsel)[rank] = (unsigned)x;
      }
    }
    RES_STAMP(4);  // reductions + block barrier + pushes
    if constexpr (GRID) {
      grid.sync();  // ghost rows (in the global images), block rows and counters have landed in L2
      RES_STAMP(5);
      if (sel) gnsel_v = __ldcg(a.gnsel + s);  // in flight while the ghost rows are pulled
      if (sel) {
        // pull this block's ghost rows of the planes just written out of its global image (L2 loads:
        // the neighbours' stores never passed through this SM's L1)
        const uint32_t *img = reinterpret_cast<const uint32_t *>(a.gimg + (size_t)rank * img_stride);
        uint32_t *pl = reinterpret_cast<uint32_t *>(smem + o_code0);
        const int nbw = nb >> 2;
        // a warp per (plane, ghost row): no index divisions, eight independent L2 loads in flight per
        // lane, so the pull costs about one L2 round trip
        for (int r12 = warp; r12 < 12; r12 += nwarp) {
          const int z = r12 >> 2, gr = r12 & 3;
          const int prow_ = (gr < 2) ? gr : nrow + gr;   // plane rows 0,1 and nrow+2, nrow+3
          const int row_at = (2 * z + (cur ^ 1)) * nbw + prow_ * W;
          for (int w0 = lane; w0 < W; w0 += 32 * 8) {
            uint32_t tmp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (w0 + 32 * u < W) tmp[u] = __ldcg(img + row_at + w0 + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (w0 + 32 * u < W) pl[row_at + w0 + 32 * u] = tmp[u];
          }
        }
        __syncthreads();  // the ghost rows are in place before the next iteration reads them
      }
      if ((s % RES_RING) == RES_RING - 1) {
        grid_fold(gfold_first, s);
        gfold_first = s + 1;
      }
    } else {
      cluster.sync();  // ghost rows, partial rows and counters of every block have landed
      RES_STAMP(5);  // cluster barrier B
    }

    if (sel) {
      // uniform lattice after the action just chosen -> the next iteration breaks (spgg.py:405)
      long long tot = 0;
      if constexpr (GRID) {
        tot = gnsel_v;
      } else {
        for (int k = 0; k < CS; ++k) tot += s_nsel[k];
      }
      if (tot == 0 || tot == n_sites) stop = j + 1;
      cur ^= 1;
    }
    last_s = s;
    if constexpr (!GRID) { pend_s = s; pend_gm = gm; }  // the row is folded inside the next barrier window
    if (tid <= RES_NI) s_redi[tid] = 0u;  // next use is behind the next block barrier
    RES_STAMP(6);  // early-exit test, fold and statistics row (rank 0)
  }

  if constexpr (GRID) {
    if (last_s >= gfold_first) grid_fold(gfold_first, last_s);   // visible since the last grid barrier
  } else if (rank == 0 && warp == 0) {
    if (pend_s >= 0) do_fold(pend_s, pend_gm);
    flush_ring();
  }

  // ---- write the state back: Q, reputation, strategy bits (ghost cells included, so the
  // per-iteration kernels can continue from these planes)
  if (rank == 0 && tid == 0) a.stop_at[rep] = stop;
#ifdef SPGG_RES_TRACE
  if (tid == 0 && a.trace)
    for (int z = 0; z < 8; ++z) a.trace[(long long)blockIdx.x * 8 + z] = trace_acc[z];
#endif
  {
    const uint8_t *Rc = smem + o_R0 + cur * nb;
    const uint8_t *Cc = smem + o_C0 + cur * nb;
    const float *sQf = reinterpret_cast<const float *>(sQ);
    for (int e = tid; e < nrow * L; e += nthr) {
      const int rr = e / L, cc = e % L;
      const int o = (rr * QR + (cc >> 2)) * 4 + (cc & 3);
      Qg[(long long)(row_start + rr) * L + cc] =
          make_float4(sQf[o], sQf[(q_plane * 4) + o], sQf[(q_plane * 8) + o], sQf[(q_plane * 12) + o]);
      store_cell<int8_t>(Rg, g, row_start + rr, cc, (int8_t)Rc[(rr + 2) * pitch + RG + cc]);
    }
    const int nW = (L + 31) >> 5;
    for (int e = tid; e < nrow * nW; e += nthr) {
      const int rr = e / nW, wi = e % nW;
      uint32_t word = 0;
      const int c0 = wi * 32, n = min(32, L - c0);
      for (int b = 0; b < n; ++b) word |= (uint32_t)(Cc[(rr + 2) * pitch + RG + c0 + b] ^ 1u) << b;
      store_bits_word(Sg, g, row_start + rr, wi, word);
    }
  }
}

// Raw statistics rows of a resident chunk -> the public row layout (include/spgg.h), with the row
// arithmetic of k_step's fold.  One thread per row; row 0 of a chunk belongs to the select-only
// first trip and carries only the reputation sum.
__global__ void k_resident_rows(const RepConst *rcs, double *stats, int n_rep, int cap, int n_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rep * n_rows) return;
  const int rep = i / n_rows, r = i - rep * n_rows;
  const RepConst &rc = rcs[rep];
  double *row = stats + ((long long)rep * cap + r) * NSTAT;
  double f[RES_NRED];
#pragma unroll
  for (int z = 0; z < RES_NRED; ++z) f[z] = row[z];
  const double gm = row[RES_RAW_GMAX];
  double s[NSTAT];
#pragma unroll
  for (int z = 0; z < NSTAT; ++z) s[z] = 0.0;
  s[ST_SUM_R] = f[17] * rc.rq;
  if (r > 0) {
    const double nc[4] = {f[0], f[1], f[2], f[3]};
    double Pc[4];
    for (int z = 0; z < 4; ++z) {  // exact-count payoff sums per class: P = ((rc*SN/5 - 5*cost*C) - lo)/span
      const double C = (z >> 1) ? 1.0 : 0.0;
      Pc[z] = ((rc.rc * f[4 + z] / 5.0 - 5.0 * rc.cost * C * nc[z]) - rc.lo * nc[z]) / rc.span;
    }
    s[ST_NC_OLD] = nc[2] + nc[3];
    s[ST_N_CD] = nc[2];
    s[ST_N_DC] = nc[1];
    s[ST_NC_NEW] = nc[1] + nc[3];
    s[ST_SUM_P] = Pc[0] + Pc[1] + Pc[2] + Pc[3];
    s[ST_SUM_P_C] = Pc[2] + Pc[3];
    s[ST_SUM_P_D] = Pc[0] + Pc[1];
    s[ST_SUM_WP_P] = rc.wP * s[ST_SUM_P];
    s[ST_SUM_REW_C] = rc.wP * (Pc[1] + Pc[3]) + rc.wR * 0.5 * (nc[1] + nc[3]);
    s[ST_SUM_REW_D] = rc.wP * (Pc[0] + Pc[2]);
    s[ST_SUM_RATIO] = f[27];
    for (int z = 0; z < 6; ++z) s[ST_GROUP0 + z] = f[8 + z];
    for (int z = 0; z < 4; ++z) {
      s[ST_SUM_Q + z] = f[18 + z];
      s[ST_SUM_Q_C + z] = f[22 + z];
      s[ST_SUM_Q_D + z] = f[18 + z] - f[22 + z];
    }
    s[ST_SUM_NI] = f[26];
    s[ST_N_BEST_POS] = f[14];
    s[ST_N_BEST_2ND] = f[15];
    s[ST_GMAX] = gm;
  }
#pragma unroll
  for (int z = 0; z < NSTAT; ++z) row[z] = s[z];
}

}  // namespace spgg
