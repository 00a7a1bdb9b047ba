// mc.cu - Mohr-Coulomb return mapping with apex smoothing on sm_100a (FP64-pipe bound).
// reference: doc/demo/demo_plasticity_mohr_coulomb.py:474-533 (return_mapping), :555 (jacfwd through
// the loop), :574-593 (vectorised C_tang_impl and its per-call statistics).  The per-point algebra is
// in mc_core.cuh; this file is the execution scheme.
//
// The work is irregular: ~2/3 of the points are elastic (a trial-stress test and a 188-byte store),
// the rest need 2..5 Newton updates of ~2.5 kflop each.  The reference's vmapped while_loop runs every
// lane until the slowest has finished (:565-571).  Here one persistent CTA per SM runs a small
// stage scheduler so that a warp always executes ONE kind of step on 32 points that need it:
//
//   stage T    32 new points: 256-bit loads, trial stress, yield test (:421-422).  Elastic points are
//              finished on the spot (tangent = C_elas, :442-443); plastic points get a state slot in
//              shared memory and are queued for S0.
//   stage S0   first residual r(y0) and ||res0|| (:500-501)               -> queue U0
//   stage U0   first Newton update (Y0 = 0: no third-derivative term) + residual + loop test
//   stage U    Newton update with the full tangent recursion + residual + loop test      -> queue U / exit
//
// A point's state (y, Y = dy/d deps, the G-partials and residual at the current iterate: 46 doubles,
// 51 when phi != psi) stays in its shared-memory slot between stages (structure-of-arrays over slots);
// queues hold slot numbers.  Warps pick the fullest-priority queue holding >= 32 entries (U > U0 > S0 > T)
// under one CTA-wide spin lock that is held for a few dozen cycles per ~10^4-cycle stage, so all 32
// lanes do the same arithmetic although points need different numbers of iterations.  Tiles of
// MC_TILE points are handed to CTAs by a global atomic counter (plastic zones cluster in real meshes).
// Statistics (:584-591: histogram of niter, max f, max ||res||) are accumulated in shared memory and
// flushed once per CTA into the context's eo_stats record.
#include "eo_common.cuh"
#include "mc_core.cuh"
#include "tab_core.cuh"
#include "tab_handle.cuh"

#define MC_THREADS 384   // 12 warps, one CTA per SM
#define MC_NSLOTS 512    // state slots per CTA (power of two: ring-buffer arithmetic)
#define MC_TILE 4096     // points per work item of the global counter
#define MC_SIMPLE_THREADS 128
#define MC_CTR_LIST 32   // word of the context's counter block that holds the length of the plastic-point list
#ifndef MCN_PREFETCH
#define MCN_PREFETCH 512  // list entries
#endif
#define MCN_FULL_DEFAULT 28
#define MCN_MIN_ACTIVE_DEFAULT 99  // 99: never wait (a warp starts with whatever lanes it can fill)
#define MCN_FIRST_LAST_DEFAULT 2
#ifndef MCN_DEFAULT_CONFIG
#define MCN_DEFAULT_CONFIG 0  // CTA shape of the lane-class Newton kernel, see mc_launch_classes
#endif

struct mc_ptrs {
  const double* deps;
  const double* sigma_n;
  double* C_tang;
  double* sigma;
  int32_t* niter;
  double* yielding;
  double* norm_res;
  double* dlambda;
};

// max(*addr, v) for doubles of either sign with NON-RETURNING atomics (RED): non-negative doubles order like signed
// integers (and beat every negative one), negative doubles order inversely to their unsigned bit patterns (and lose
// against every non-negative one).  The issuing thread does not wait for the L2 round trip.
__device__ __forceinline__ void mc_atomic_max_f64(double* addr, double v) {
  if (!(v == v)) return;  // NaN never becomes a maximum
  if (v >= 0.0)
    atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v + 0.0));  // + 0.0: -0.0 -> +0.0
  else
    atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// warp-wide maximum of doubles (NaN lanes are ignored; -inf when every lane holds NaN) with two REDUX instructions on
// an order-preserving 64-bit key instead of a ten-shuffle butterfly
__device__ __forceinline__ double mc_warp_max_f64(double v) {
  unsigned long long key = (unsigned long long)__double_as_longlong(v);
  key = (key >> 63) ? ~key : (key | 0x8000000000000000ull);
  if (!(v == v)) key = 0x000fffffffffffffull;  // key of -inf
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  unsigned long long m = ((unsigned long long)mh << 32) | ml;
  m = (m >> 63) ? (m & 0x7fffffffffffffffull) : ~m;
  return __longlong_as_double((long long)m);
}

__device__ __forceinline__ void mc_store_aux(const mc_ptrs& P, int64_t i, int32_t niter, double yielding,
                                             double norm_res, double dlambda) {
  if (P.niter) P.niter[i] = niter;
  if (P.yielding) eo_st64(P.yielding + i, yielding);
  if (P.norm_res) eo_st64(P.norm_res + i, norm_res);
  if (P.dlambda) eo_st64(P.dlambda + i, dlambda);
}

__device__ __forceinline__ void mc_store_point(const mc_ptrs& P, int64_t i, const double Ct[16], const double sig[4]) {
  double* C = P.C_tang + 16 * i;
  eo_st256(C + 0, Ct[0], Ct[1], Ct[2], Ct[3]);
  eo_st256(C + 4, Ct[4], Ct[5], Ct[6], Ct[7]);
  eo_st256(C + 8, Ct[8], Ct[9], Ct[10], Ct[11]);
  eo_st256(C + 12, Ct[12], Ct[13], Ct[14], Ct[15]);
  eo_st256(P.sigma + 4 * i, sig[0], sig[1], sig[2], sig[3]);
}

enum { MC_Q_S0 = 0, MC_Q_U0 = 1, MC_Q_U = 2, MC_Q_FREE = 3, MC_STAGE_T = 3, MC_STAGE_WAIT = 4, MC_STAGE_EXIT = 5 };
#define MC_EMPTY 0xFFFFu
#define MC_UNITS_PER_TILE (MC_TILE / 32)

// Bounded multi-producer / multi-consumer ring of slot numbers with per-cell full/empty state
// (a cell holds MC_EMPTY or a slot number).  `cnt` counts fully published entries and is what consumers
// claim from; `head`/`tail` are free-running positions.  No locks: a consumer that claimed a cell whose
// producer is still writing spins on the cell for a few cycles, and vice versa.
struct mc_queue {
  unsigned int cnt, head, tail, pad;
};

// leader lane: claim between want_min and 32 published entries; returns the number taken (0: none)
__device__ __forceinline__ int mc_q_claim(mc_queue* q, unsigned int want_min, unsigned int& base) {
  for (int tries = 0; tries < 4; ++tries) {
    const unsigned int c = *(volatile unsigned int*)&q->cnt;
    if (c < want_min || c == 0) return 0;
    const unsigned int t = c < 32u ? c : 32u;
    if (atomicCAS(&q->cnt, c, c - t) == c) {
      base = atomicAdd(&q->head, t);
      return (int)t;
    }
  }
  return 0;
}

__device__ __forceinline__ int mc_q_read(volatile unsigned short* ring, unsigned int pos) {
  unsigned short v;
  while ((v = ring[pos & (MC_NSLOTS - 1)]) == MC_EMPTY) {
  }
  ring[pos & (MC_NSLOTS - 1)] = MC_EMPTY;
  return (int)v;
}

// warp-collective: lanes with slot >= 0 append it
__device__ __forceinline__ void mc_q_push(mc_queue* q, volatile unsigned short* ring, int slot, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, slot >= 0);
  if (m == 0) return;
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(&q->tail, (unsigned)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (slot >= 0) {
    const unsigned int pos = (base + __popc(m & ((1u << lane) - 1u))) & (MC_NSLOTS - 1);
    while (ring[pos] != MC_EMPTY) {
    }
    ring[pos] = (unsigned short)slot;
  }
  __threadfence_block();
  __syncwarp();
  if (lane == 0) atomicAdd(&q->cnt, (unsigned)__popc(m));
}

// LISTED: the yield test has already been done by mc_trial_kernel (two-pass scheme); the T stage then only fetches
// plastic points from the list (point index + f(trial)) and opens their slots.
template <bool ASSOC, int AFF, bool LISTED>
__global__ void __launch_bounds__(MC_THREADS, 1) mc_kernel(const mc_consts k, const mc_ptrs P, const int64_t n_in,
                                                           eo_stats* __restrict__ stats, unsigned int* tile_ctr,
                                                           const int32_t* __restrict__ list,
                                                           const double* __restrict__ list_yl) {
  const int64_t n = LISTED ? (int64_t)tile_ctr[MC_CTR_LIST] : n_in;
  extern __shared__ double s_slots[];  // [ASSOC ? MC_NF_ASSOC : MC_NF][MC_NSLOTS]
  __shared__ unsigned short s_ring[4][MC_NSLOTS];
  __shared__ long long s_pt[MC_NSLOTS];
  __shared__ int s_it[MC_NSLOTS];
  __shared__ mc_queue s_q[4];
  __shared__ unsigned long long s_work;  // (tile index << 16) | next 32-point unit within the tile
  __shared__ int s_inflight, s_done, s_fetching;  // inflight: plastic points in the system + T stages running
  __shared__ unsigned int s_hist[EO_NITER_BINS];
  __shared__ unsigned int s_nonconv, s_nonfinite, s_plastic;

  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t ntiles = (n + MC_TILE - 1) / MC_TILE;
  for (int b = tid; b < EO_NITER_BINS; b += MC_THREADS) s_hist[b] = 0;
  for (int b = tid; b < MC_NSLOTS; b += MC_THREADS) {
    s_ring[MC_Q_FREE][b] = (unsigned short)b;
    s_ring[MC_Q_S0][b] = s_ring[MC_Q_U0][b] = s_ring[MC_Q_U][b] = MC_EMPTY;
  }
  if (tid == 0) {
    s_nonconv = s_nonfinite = s_plastic = 0;
    s_inflight = 0, s_done = 0, s_fetching = 0;
    for (int q = 0; q < 4; ++q) s_q[q].cnt = 0, s_q[q].head = 0, s_q[q].tail = 0;
    s_q[MC_Q_FREE].cnt = MC_NSLOTS;
    s_work = 0xFFFFull;  // "exhausted": the first warp to look fetches a tile
  }
  __syncthreads();
  double max_f = -INFINITY, max_res = 0.0;
  int max_it = 0;

  for (;;) {
    // ------------------------------------------------------------------ pick a stage (lane 0, lock-free)
    int stage = MC_STAGE_WAIT, take = 0;
    unsigned int qpos = 0;
    long long in0 = 0;
    if (lane == 0) {
      const unsigned int cU = *(volatile unsigned int*)&s_q[MC_Q_U].cnt, cU0 = *(volatile unsigned int*)&s_q[MC_Q_U0].cnt,
                         cS0 = *(volatile unsigned int*)&s_q[MC_Q_S0].cnt, cF = *(volatile unsigned int*)&s_q[MC_Q_FREE].cnt;
      unsigned long long w = *(volatile unsigned long long*)&s_work;
      long long tile = (long long)(w >> 16);
      long long units = (tile < ntiles) ? (((n - tile * MC_TILE < MC_TILE ? n - tile * MC_TILE : MC_TILE) + 31) / 32) : 0;
      bool inputs = (long long)(w & 0xFFFFull) < units;
      if (!inputs && !*(volatile int*)&s_done && atomicCAS(&s_fetching, 0, 1) == 0) {
        // fetch the next tile (one warp at a time; re-check under the flag)
        w = *(volatile unsigned long long*)&s_work;
        tile = (long long)(w >> 16);
        units = (tile < ntiles) ? (((n - tile * MC_TILE < MC_TILE ? n - tile * MC_TILE : MC_TILE) + 31) / 32) : 0;
        if (!((long long)(w & 0xFFFFull) < units)) {
          const unsigned int t = atomicAdd(tile_ctr, 1u);
          if ((int64_t)t >= ntiles) {
            *(volatile int*)&s_done = 1;
          } else {
            atomicExch(&s_work, (unsigned long long)t << 16);
            inputs = true;
          }
        } else {
          inputs = true;
        }
        __threadfence_block();
        atomicExch(&s_fetching, 0);
      }
      int want = -1;
      if (AFF) {
        // stage affinity per SM sub-partition (warp w runs on sub-partition w & 3): keeping one kind of stage
        // on a sub-partition keeps that stage's code in its instruction cache; falls through to the priority
        // rule when the preferred queue cannot fill a warp
        const int sp = (tid >> 5) & 3;
        if (sp == 0) {
          if (inputs && cF >= 32 && cU < 192) want = MC_STAGE_T;
          else if (cS0 >= 32) want = MC_Q_S0;
        } else if (sp == 1) {
          if (cU0 >= 32) want = MC_Q_U0;
          else if (cS0 >= 32) want = MC_Q_S0;
        } else if (cU >= 32) want = MC_Q_U;
      }
      if (want >= 0) {
      } else if (cU >= 32) want = MC_Q_U;
      else if (cU0 >= 32) want = MC_Q_U0;
      else if (cS0 >= 32) want = MC_Q_S0;
      else if (inputs && cF >= 32) want = MC_STAGE_T;
      else if (cU | cU0 | cS0) want = (cU >= cU0 && cU >= cS0) ? MC_Q_U : (cU0 >= cS0 ? MC_Q_U0 : MC_Q_S0);
      if (want >= 0 && want <= MC_Q_U) {
        take = mc_q_claim(&s_q[want], 1, qpos);
        if (take > 0) stage = want;
      } else if (want == MC_STAGE_T) {
        take = mc_q_claim(&s_q[MC_Q_FREE], 32, qpos);  // reserve 32 slots; unused ones go back after the stage
        if (take == 32) {
          atomicAdd(&s_inflight, 1);  // before the claim: "inflight == 0 and no inputs" then means finished
          const unsigned long long w2 = atomicAdd(&s_work, 1ull);
          const long long tile2 = (long long)(w2 >> 16), off = (long long)(w2 & 0xFFFFull);
          const long long pts = tile2 < ntiles ? (n - tile2 * MC_TILE < MC_TILE ? n - tile2 * MC_TILE : MC_TILE) : 0;
          if (off * 32 < pts) {
            stage = MC_STAGE_T;
            in0 = tile2 * MC_TILE + off * 32;
            take = (int)(pts - off * 32 < 32 ? pts - off * 32 : 32);
          } else {
            stage = MC_STAGE_T;  // lost the race for the last unit: run an empty T stage that returns the slots
            take = 0;
          }
        } else {
          take = 0;
        }
      }
      if (stage == MC_STAGE_WAIT && *(volatile int*)&s_done && *(volatile int*)&s_inflight == 0) {
        // nothing claimed, no tile left, no plastic point and no T stage in flight anywhere in the CTA
        const unsigned long long w3 = *(volatile unsigned long long*)&s_work;
        const long long tile3 = (long long)(w3 >> 16);
        const long long pts3 = tile3 < ntiles ? (n - tile3 * MC_TILE < MC_TILE ? n - tile3 * MC_TILE : MC_TILE) : 0;
        if (!((long long)(w3 & 0xFFFFull) * 32 < pts3)) stage = MC_STAGE_EXIT;
      }
    }
    stage = __shfl_sync(0xffffffffu, stage, 0);
    if (stage == MC_STAGE_EXIT) break;
#ifdef MC_DEBUG_COUNTERS
    if (lane == 0) {
      if (stage == MC_STAGE_WAIT) atomicAdd(tile_ctr + 9, 1u);
      else {
        atomicAdd(tile_ctr + 1 + 2 * (stage == MC_STAGE_T ? 0 : stage + 1), 1u);
        atomicAdd(tile_ctr + 2 + 2 * (stage == MC_STAGE_T ? 0 : stage + 1), (unsigned)take);
      }
    }
#endif
    if (stage == MC_STAGE_WAIT) {
      __nanosleep(100);
      continue;
    }
#ifdef MC_DEBUG_COUNTERS
    const long long dbg_t0 = clock64();
#endif
    take = __shfl_sync(0xffffffffu, take, 0);
    qpos = __shfl_sync(0xffffffffu, qpos, 0);

    int slotA = -1;  // slot this lane returns to the FREE list after the stage
    int slotB = -1;  // slot this lane hands on to the next queue
    if (stage == MC_STAGE_T) {
      // ---------------------------------------------------------------- stage T
      in0 = __shfl_sync(0xffffffffu, in0, 0);
      const int myres = mc_q_read(s_ring[MC_Q_FREE], qpos + lane);  // the lane-th reserved slot
      bool plastic = false;
      double yl = 0.0, sn[4], Cde[4];
      int64_t i = in0 + lane;
      if (LISTED) {
        if (lane < take) {
          yl = list_yl[i];
          i = list[i];
          const eo_d4 e = eo_ld256(P.deps + 4 * i);
          const eo_d4 sg = eo_ld256(P.sigma_n + 4 * i);
          const double de[4] = {e.x, e.y, e.z, e.w};
          sn[0] = sg.x, sn[1] = sg.y, sn[2] = sg.z, sn[3] = sg.w;
          mc_Cmul(k, de, Cde);
          plastic = true;
        }
      } else if (lane < take) {
        const eo_d4 e = eo_ld256(P.deps + 4 * i);
        const eo_d4 sg = eo_ld256(P.sigma_n + 4 * i);
        const double de[4] = {e.x, e.y, e.z, e.w};
        sn[0] = sg.x, sn[1] = sg.y, sn[2] = sg.z, sn[3] = sg.w;
        yl = mc_trial(k, de, sn, Cde);
        max_f = fmax(max_f, yl);
        if (yl <= 0.0) {
          double sig[4], Ct[16], nr, dl;
          const int32_t it = mc_elastic(k, sn, Cde, sig, Ct, nr, dl);
          mc_store_point(P, i, Ct, sig);
          mc_store_aux(P, i, it, yl, nr, dl);
          max_res = fmax(max_res, nr);
          max_it = max(max_it, it);
          atomicAdd(&s_hist[it], 1u);
          if (!(isfinite(sig[0]) && isfinite(sig[1]) && isfinite(sig[2]) && isfinite(sig[3])))
            atomicAdd(&s_nonfinite, 1u);
        } else {
          plastic = true;  // a NaN predicate takes the plastic branch, like `yielding <= 0.0` being false
        }
      }
      const unsigned pm = __ballot_sync(0xffffffffu, plastic);
      const int rank = __popc(pm & ((1u << lane) - 1u));
      const int npl = __popc(pm);
      // the plastic lane of rank r takes reserved slot r; reserved slots >= npl go back
      const int got = __shfl_sync(0xffffffffu, myres, plastic ? rank : 0);
      if (plastic) {
        slotB = got;
        const mc_slot sl{s_slots + slotB, MC_NSLOTS};
        mc_slot_init(sl, sn, Cde, yl);
        s_pt[slotB] = i;
        s_it[slotB] = 0;
      }
      if (lane >= npl) slotA = myres;
      if (lane == 0) atomicAdd(&s_inflight, npl - 1);  // the points enter before this T stage leaves
    } else {
      // ---------------------------------------------------------------- stages S0 / U0 / U
      bool fin = false, nonconv = false, nonfin = false;
      int bin = 0;
      if (lane < take) {
        const int slot = mc_q_read(s_ring[stage], qpos + lane);
        const mc_slot sl{s_slots + slot, MC_NSLOTS};
        int32_t it = s_it[slot];
        const bool out = mc_stage<ASSOC>(k, stage, sl, it);
        s_it[slot] = it;
        if (out) {
          const int64_t i = s_pt[slot];
          double Ct[16], sig[4], nr, dl;
          mc_slot_result(sl, it, Ct, sig, nr, dl);
          mc_store_point(P, i, Ct, sig);
          mc_store_aux(P, i, it, sl[MC_F_YIELD], nr, dl);
          max_res = fmax(max_res, nr);
          max_it = max(max_it, it);
          fin = true;
          bin = min(it, EO_NITER_BINS - 1);
          nonconv = it >= k.nitermax;
          nonfin = !(isfinite(sig[0]) && isfinite(sig[1]) && isfinite(sig[2]) && isfinite(sig[3]));
          slotA = slot;
        } else {
          slotB = slot;
        }
      }
      // statistics of the points that left the loop: warp-aggregated (one shared-memory atomic per counter and per
      // distinct iteration count instead of one per lane)
      const unsigned fm = __ballot_sync(0xffffffffu, fin);
      if (fm) {
        const unsigned ncm = __ballot_sync(0xffffffffu, nonconv), nfm = __ballot_sync(0xffffffffu, nonfin);
        if (fin) {
          const unsigned grp = __match_any_sync(fm, bin);
          if (lane == __ffs(grp) - 1) atomicAdd(&s_hist[bin], (unsigned)__popc(grp));
        }
        if (lane == 0) {
          atomicAdd(&s_plastic, (unsigned)__popc(fm));
          atomicAdd(&s_inflight, -__popc(fm));
          if (ncm) atomicAdd(&s_nonconv, (unsigned)__popc(ncm));
          if (nfm) atomicAdd(&s_nonfinite, (unsigned)__popc(nfm));
        }
      }
    }

    // ------------------------------------------------------------------ hand the slots on
    const int qB = stage == MC_STAGE_T ? MC_Q_S0 : (stage == MC_Q_S0 ? MC_Q_U0 : MC_Q_U);
    __syncwarp();
    mc_q_push(&s_q[qB], s_ring[qB], slotB, lane);
    mc_q_push(&s_q[MC_Q_FREE], s_ring[MC_Q_FREE], slotA, lane);
    __syncwarp();
#ifdef MC_DEBUG_COUNTERS
    if (lane == 0)  // warp-cycles per stage kind: 64-bit counters at words 16.. (T, S0, U0, U)
      atomicAdd(reinterpret_cast<unsigned long long*>(tile_ctr + 16) + (stage == MC_STAGE_T ? 0 : stage + 1),
                (unsigned long long)(clock64() - dbg_t0));
#endif
  }

  // ------------------------------------------------------------------ statistics flush
  __syncthreads();
  for (int b = tid; b < EO_NITER_BINS; b += MC_THREADS)
    if (s_hist[b]) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->niter_hist[b]), (unsigned long long)s_hist[b]);
  if (tid == 0) {
    if (s_plastic) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_plastic), (unsigned long long)s_plastic);
    if (s_nonconv) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_nonconverged), (unsigned long long)s_nonconv);
    if (s_nonfinite) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_nonfinite), (unsigned long long)s_nonfinite);
    if (blockIdx.x == 0 && !LISTED) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)n);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    max_f = fmax(max_f, __shfl_xor_sync(0xffffffffu, max_f, o));
    max_res = fmax(max_res, __shfl_xor_sync(0xffffffffu, max_res, o));
    max_it = max(max_it, __shfl_xor_sync(0xffffffffu, max_it, o));
  }
  if (lane == 0) {
    mc_atomic_max_f64(&stats->f_max, max_f);
    mc_atomic_max_f64(&stats->res_max, max_res);
    mc_atomic_max_f64(&stats->niter_max, (double)max_it);
  }
}

// Two-pass scheme, pass 1: one thread per point at full occupancy (HBM-bound for elastic points): trial stress and
// yield test (:421-422); elastic points are finished here, plastic points are appended (warp-aggregated) to a list
// that mc_kernel<.., LISTED> works off.  The long dependent FP64 chain of the yield test (asin, sincos, sqrt) is
// latency-bound at the 12 warps/SM of the persistent kernel; here 8x more warps hide it.
// NB > 0: the strain increment is not read but TABULATED here - Mandel strain of the P1/P2/P3 vector field `src.u` at this
// thread's point (cell = i / nq), like tab_vm_kernel - and, for plastic points only, stored next to the list entry
// (`list_deps`), where pass 2 picks it up: the strain array of the whole mesh is never written (eo_mc_eval_tabulated).
struct mc_tab_src {
  const int32_t* dofmap;
  const int32_t* x_dofmap;
  const double* x;
  const double* u;
};

template <bool ASSOC, int NB>
__global__ void __launch_bounds__(256, NB > 0 ? 3 : 4) mc_trial_kernel(const mc_consts k, const mc_ptrs P, const int64_t n,
                                                       eo_stats* __restrict__ stats, unsigned int* ctr,
                                                       int32_t* __restrict__ list, double* __restrict__ list_yl,
                                                       const tab_tables* __restrict__ T, const mc_tab_src src,
                                                       double* __restrict__ list_deps) {
  __shared__ unsigned int s_hist0, s_hist1, s_histx, s_nonfinite;
  __shared__ double s_dphi[NB > 0 ? EO_TAB_MAX_NQ : 1][2][NB > 0 ? NB : 1];
  __shared__ double s_dpsi[2][3];
  if (NB > 0) {
    for (int t = threadIdx.x; t < T->nq * 2 * NB; t += blockDim.x) {
      const int q = t / (2 * NB), kk = (t / NB) % 2, a = t % NB;
      s_dphi[q][kk][a] = T->dphi[kk][q][a];
    }
    if (threadIdx.x < 6) s_dpsi[threadIdx.x / 3][threadIdx.x % 3] = T->dpsi[threadIdx.x / 3][threadIdx.x % 3];
  }
  if (threadIdx.x == 0) s_hist0 = s_hist1 = s_histx = s_nonfinite = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  double max_f = -INFINITY, max_res = 0.0;
  int max_it = 0;
  bool plastic = false;
  double yl = 0.0;
  double de[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < n) {
    const eo_d4 sg = eo_ld256(P.sigma_n + 4 * i);
    if (NB > 0) {
      const int nq = T->nq;
      const int64_t c = i / nq;
      const int q = int(i - c * nq);
      double w[NB > 0 ? NB : 1][2], xv[3][2], J[2][2], K[2][2];
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(src.u) + __ldg(src.dofmap + c * NB + a));
        w[a][0] = v.x, w[a][1] = v.y;
      }
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int64_t node = __ldg(src.x_dofmap + c * 3 + v);
        xv[v][0] = __ldg(src.x + 3 * node), xv[v][1] = __ldg(src.x + 3 * node + 1);
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) J[a][b] = xv[0][a] * s_dpsi[b][0] + xv[1][a] * s_dpsi[b][1] + xv[2][a] * s_dpsi[b][2];
      tab_inverse<2>(J, K);
      double G[2][2], grad[2][2], val[2] = {0.0, 0.0};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < NB; ++a) acc += w[a][cc] * s_dphi[q][kk][a];
          G[cc][kk] = acc;
        }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
#pragma unroll
        for (int j = 0; j < 2; ++j) grad[cc][j] = G[cc][0] * K[0][j] + G[cc][1] * K[1][j];
      tab_operand<2, 2>(2, val, grad, de);
    } else {
      const eo_d4 e = eo_ld256(P.deps + 4 * i);
      de[0] = e.x, de[1] = e.y, de[2] = e.z, de[3] = e.w;
    }
    const double sn[4] = {sg.x, sg.y, sg.z, sg.w};
    double Cde[4];
    yl = mc_trial(k, de, sn, Cde);
    max_f = yl;
    if (yl <= 0.0) {
      double sig[4], Ct[16], nr, dl;
      const int32_t it = mc_elastic(k, sn, Cde, sig, Ct, nr, dl);
      mc_store_point(P, i, Ct, sig);
      mc_store_aux(P, i, it, yl, nr, dl);
      max_res = nr;
      max_it = it;
      if (it == 0) atomicAdd(&s_hist0, 1u);
      else if (it == 1) atomicAdd(&s_hist1, 1u);
      else atomicAdd(reinterpret_cast<unsigned long long*>(&stats->niter_hist[min(it, EO_NITER_BINS - 1)]), 1ull);
      if (!(isfinite(sig[0]) && isfinite(sig[1]) && isfinite(sig[2]) && isfinite(sig[3]))) atomicAdd(&s_nonfinite, 1u);
    } else {
      plastic = true;  // NaN predicate -> plastic branch, like `yielding <= 0.0` being false
      // the aux arrays are written for EVERY point here (pass 2 overwrites the plastic entries): a 32-byte sector written
      // in part costs a read-fill from HBM when it leaves L2
      mc_store_aux(P, i, 0, yl, 0.0, 0.0);
    }
  }
  // ---- CTA-aggregated epilogue: ONE list reservation per CTA (a per-warp atomicAdd on the single list counter would
  //      serialise 6 x 10^5 same-address atomics per 2 x 10^7 points).  Only that one returning atomic sits between the
  //      two barriers; the statistics go out afterwards as non-returning reductions.
  __shared__ unsigned int s_wbase[8];
  __shared__ double s_mf[8], s_mr[8];
  __shared__ int s_mi[8];
  const int warp = threadIdx.x >> 5;
  const unsigned pm = __ballot_sync(0xffffffffu, plastic);
  max_f = mc_warp_max_f64(max_f);
  if (__any_sync(0xffffffffu, max_res != 0.0)) max_res = mc_warp_max_f64(max_res);  // exactly 0 for elastic points
  max_it = __reduce_max_sync(0xffffffffu, max_it);
  if (lane == 0) s_wbase[warp] = __popc(pm), s_mf[warp] = max_f, s_mr[warp] = max_res, s_mi[warp] = max_it;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const unsigned int c = s_wbase[w];
      s_wbase[w] = total;
      total += c;
    }
    const unsigned int base = total ? atomicAdd(ctr + MC_CTR_LIST, total) : 0u;
#pragma unroll
    for (int w = 0; w < 8; ++w) s_wbase[w] += base;
  }
  __syncthreads();
  if (plastic) {
    const unsigned int pos = s_wbase[warp] + __popc(pm & ((1u << lane) - 1u));
    list[pos] = (int32_t)i;
    list_yl[pos] = yl;
    if (NB > 0) eo_st256(list_deps + 4 * size_t(pos), de[0], de[1], de[2], de[3]);
  }
  if (threadIdx.x == 32) {  // a thread of another warp than the one that reserved the list space
    // the maxima saturate after a few CTAs: look first (three independent loads), reduce only what raises a maximum
    const double cf = *(volatile double*)&stats->f_max, cr = *(volatile double*)&stats->res_max,
                 ci = *(volatile double*)&stats->niter_max;
    double mf = -INFINITY, mr = 0.0;
    int mi = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) mf = fmax(mf, s_mf[w]), mr = fmax(mr, s_mr[w]), mi = max(mi, s_mi[w]);
    if (!(mf <= cf)) mc_atomic_max_f64(&stats->f_max, mf);
    if (!(mr <= cr)) mc_atomic_max_f64(&stats->res_max, mr);
    if (!((double)mi <= ci)) mc_atomic_max_f64(&stats->niter_max, (double)mi);
    if (s_hist0) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->niter_hist[0]), (unsigned long long)s_hist0);
    if (s_hist1) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->niter_hist[1]), (unsigned long long)s_hist1);
    if (s_nonfinite) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_nonfinite), (unsigned long long)s_nonfinite);
    if (blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)n);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-pass scheme, pass 2 (default): persistent CTAs work off the plastic-point list with LANE-CLASS slots.
//
// A point's Newton state lives in a shared-memory slot for its whole life, like in mc_kernel, but slot (b, l) =
// b * 32 + l belongs to lane class l: only lane l of whichever warp ever touches it.  Field f of slot (b, l) sits at
// s_slots[f * NSLOT + b * 32 + l], so the 32 lanes of a warp always access 32 consecutive doubles - no shared-memory
// bank conflicts by construction (the ring scheduler hands arbitrary slots to lanes: 58 % of its shared-memory
// wavefronts are conflict replays, profiles/r2_mc_ncu_summary.md).  Scheduling state is two words per class:
// s_free[l] / s_ready[l], bit b = slot (b, l) is unused / waits for its next Newton update.  A lane allocates and
// hands on slots with one atomicAnd / atomicOr on its own class word (32 consecutive words per warp instruction); there
// are no queues, no ring positions and no leader-lane lock.
//
// Stages (one inlined copy of mc_stage serves both):
//   F  first visit of up to 32 listed points: load, first residual, first update, residual, loop test (kinds 0 + 1) and -
//      by default - one full update more (kind 2): 63 % of the plastic points of the demo's stress paths converge with
//      their second update and so never leave the warp that loaded them
//   U  one further Newton update with the full tangent recursion + residual + loop test                (kind 2)
// A warp takes U when at least MCN_FULL classes have a waiting point, else F when that many classes have a free slot
// and list entries remain, else whichever fills more lanes.  Points that left the loop are written out from the slot.
struct mcn_policy {
  int full;         // lanes a stage must fill to start at once
  int min_active;   // ... unless fewer than this many warps of the CTA are inside a stage
  int first_last;   // last stage kind of a point's FIRST visit: 0 = first residual only, 1 = + first update, 2 = + second update
  unsigned sleep_ns;
};

template <bool ASSOC, int NWARPS, int DEPTH>
__global__ void __launch_bounds__(NWARPS * 32, 1) mc_newton_kernel(const mc_consts k, const mc_ptrs P, eo_stats* __restrict__ stats,
                                                                   unsigned int* ctr, const int32_t* __restrict__ list,
                                                                   const double* __restrict__ list_yl, const mcn_policy pol,
                                                                   const double* __restrict__ list_deps) {
  constexpr int NSLOT = 32 * DEPTH;
  constexpr unsigned ALL = DEPTH == 32 ? 0xffffffffu : ((1u << DEPTH) - 1u);
  const int64_t n = (int64_t)ctr[MC_CTR_LIST];
  extern __shared__ double s_slots[];  // [ASSOC ? MC_NF_ASSOC : MC_NF][NSLOT]
  __shared__ unsigned int s_free[32], s_ready[32];
  __shared__ int s_pt[NSLOT];
  __shared__ int s_it[NSLOT];
  __shared__ unsigned long long s_work;  // (tile index << 16) | next list entry within the tile
  __shared__ int s_inflight, s_done, s_fetching, s_active;  // s_active: warps inside a stage
  __shared__ unsigned int s_hist[EO_NITER_BINS];
  __shared__ unsigned int s_nonconv, s_nonfinite, s_plastic;

  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t ntiles = (n + MC_TILE - 1) / MC_TILE;
  for (int b = tid; b < EO_NITER_BINS; b += NWARPS * 32) s_hist[b] = 0;
  if (tid < 32) s_free[tid] = ALL, s_ready[tid] = 0;
  if (tid == 0) {
    s_nonconv = s_nonfinite = s_plastic = 0;
    s_inflight = 0, s_done = 0, s_fetching = 0, s_active = 0;
    s_work = 0xFFFFull;  // "exhausted": the first warp to look fetches a tile
  }
  __syncthreads();
  double max_res = 0.0;
  int max_it = 0;
  auto tile_points = [&](long long tile) -> long long {
    return tile < ntiles ? (n - tile * MC_TILE < MC_TILE ? n - tile * MC_TILE : MC_TILE) : 0;
  };

  for (;;) {
    // ------------------------------------------------------------------ what could this warp do?
    const unsigned rdy = *(volatile unsigned int*)&s_ready[lane], fre = *(volatile unsigned int*)&s_free[lane];
    const int nU = __popc(__ballot_sync(0xffffffffu, rdy != 0)), nF = __popc(__ballot_sync(0xffffffffu, fre != 0));
    int stage = MC_STAGE_WAIT;
    if (lane == 0) {
      unsigned long long w = *(volatile unsigned long long*)&s_work;
      bool inputs = (long long)(w & 0xFFFFull) < tile_points((long long)(w >> 16));
      if (!inputs && !*(volatile int*)&s_done && nF > 0 && atomicCAS(&s_fetching, 0, 1) == 0) {
        w = *(volatile unsigned long long*)&s_work;  // one warp at a time fetches the next tile; re-check under the flag
        if (!((long long)(w & 0xFFFFull) < tile_points((long long)(w >> 16)))) {
          const unsigned int t = atomicAdd(ctr, 1u);
          if ((int64_t)t >= ntiles) {
            *(volatile int*)&s_done = 1;
          } else {
            atomicExch(&s_work, (unsigned long long)t << 16);
            inputs = true;
          }
        } else {
          inputs = true;
        }
        __threadfence_block();
        atomicExch(&s_fetching, 0);
      }
      const int f = inputs ? nF : 0;
      // a partly filled warp occupies the FP64 pipe like a full one: while enough other warps are inside a stage (they
      // will hand on / free slots within microseconds) it is cheaper to look again than to start with idle lanes
      const bool patient = *(volatile int*)&s_active >= pol.min_active;
      if (nU >= pol.full) stage = MC_Q_U;
      else if (f >= pol.full) stage = MC_STAGE_T;
      else if (patient && (nU | f)) stage = MC_STAGE_WAIT;
      else if (nU > 0 && nU >= f) stage = MC_Q_U;
      else if (f > 0) stage = MC_STAGE_T;
      else if (*(volatile int*)&s_done && *(volatile int*)&s_inflight == 0) {
        const unsigned long long w3 = *(volatile unsigned long long*)&s_work;
        if (!((long long)(w3 & 0xFFFFull) < tile_points((long long)(w3 >> 16)))) stage = MC_STAGE_EXIT;
      }
    }
    stage = __shfl_sync(0xffffffffu, stage, 0);
    if (stage == MC_STAGE_EXIT) break;
    if (stage == MC_STAGE_WAIT) {
      __nanosleep(pol.sleep_ns);
      continue;
    }
    if (lane == 0) atomicAdd(&s_active, 1);

    // ------------------------------------------------------------------ take a slot of this lane's class
    int slot = -1;
    {
      volatile unsigned int* word = stage == MC_Q_U ? &s_ready[lane] : &s_free[lane];
      for (;;) {
        const unsigned m = *word;
        if (m == 0) break;
        const unsigned bit = m & (0u - m);
        if (atomicAnd(const_cast<unsigned int*>(word), ~bit) & bit) {
          slot = (__ffs(bit) - 1) * 32 + lane;
          break;
        }
      }
    }
    // acquire side of the hand-over (the release side is the fence before the atomicOr on s_ready below): the slot's
    // fields are read only after the claim.  compute-sanitizer's racecheck models barriers only and still lists these
    // slot accesses as hazards (profiles/r2_sanitizer.md).
    __threadfence_block();
    int kind0 = 2, kind1 = 2;
    if (stage == MC_STAGE_T) {
      // ---------------------------------------------------------------- stage F: claim list entries for the lanes that
      //                                                                  hold a slot, load the points, open the slots
      kind0 = 0, kind1 = pol.first_last;
      const unsigned sm = __ballot_sync(0xffffffffu, slot >= 0);
      const int want = __popc(sm);
      long long e0 = 0;
      int got = 0, ahead = 0;  // ahead: entries of the tile from this claim on
      if (lane == 0 && want > 0) {
        atomicAdd(&s_inflight, want);  // before the claim: "done and inflight == 0" then means finished
        const unsigned long long w2 = atomicAdd(&s_work, (unsigned long long)want);
        const long long tile2 = (long long)(w2 >> 16), off = (long long)(w2 & 0xFFFFull), pts = tile_points(tile2);
        if (off < pts) {
          e0 = tile2 * MC_TILE + off;
          got = (int)(pts - off < want ? pts - off : want);
          ahead = (int)(pts - off);
        }
        if (got < want) atomicAdd(&s_inflight, got - want);
      }
      e0 = __shfl_sync(0xffffffffu, e0, 0);
      got = __shfl_sync(0xffffffffu, got, 0);
      ahead = __shfl_sync(0xffffffffu, ahead, 0);
      const int rank = __popc(sm & ((1u << lane) - 1u));
      if (slot >= 0 && rank >= got) {  // no list entry for this slot: hand it back
        atomicOr(&s_free[lane], 1u << (slot >> 5));
        slot = -1;
      }
      if (slot >= 0) {
        const int64_t e = e0 + rank;
        const double yl = list_yl[e];
        const int i = list[e];
        // the inputs of the entry MCN_PREFETCH places further on (claimed about one round of the CTA's warps from now)
        // are pulled into L2 while this point is worked on: a first visit then waits for an L2 hit, not for HBM
        // (measured: 12.58 -> 12.39 ms per 1e8 points)
        const int ip = rank + MCN_PREFETCH < ahead ? list[e + MCN_PREFETCH] : -1;
        const eo_d4 de4 = list_deps ? eo_ld256(list_deps + 4 * e) : eo_ld256(P.deps + 4 * (int64_t)i);
        const eo_d4 sg = eo_ld256(P.sigma_n + 4 * (int64_t)i);
        if (ip >= 0) {
          if (!list_deps) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.deps + 4 * (int64_t)ip));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(P.sigma_n + 4 * (int64_t)ip));
        }
        const double de[4] = {de4.x, de4.y, de4.z, de4.w}, sn[4] = {sg.x, sg.y, sg.z, sg.w};
        double Cde[4];
        mc_Cmul(k, de, Cde);
        const mc_slot sl{s_slots + slot, NSLOT};
        mc_slot_init(sl, sn, Cde, yl);
        s_pt[slot] = i;
        s_it[slot] = 0;
      }
    }

    // ------------------------------------------------------------------ Newton stage(s) on the slot
    bool fin = false, nonconv = false, nonfin = false;
    int bin = 0;
    if (slot >= 0) {
      const mc_slot sl{s_slots + slot, NSLOT};
      int32_t it = s_it[slot];
      if (stage == MC_Q_U && it == 0) kind0 = kind1 = 1;  // first_last == 0: the first update runs in its own visit
      bool out = false;
#pragma unroll 1
      for (int kind = kind0;; ++kind) {
        out = mc_stage<ASSOC>(k, kind, sl, it);
        if (out || kind >= kind1) break;
      }
      if (out) {
        const int64_t i = s_pt[slot];
        double Ct[16], sig[4], nr, dl;
        mc_slot_result(sl, it, Ct, sig, nr, dl);
        mc_store_point(P, i, Ct, sig);
        mc_store_aux(P, i, it, sl[MC_F_YIELD], nr, dl);
        max_res = fmax(max_res, nr);
        max_it = max(max_it, it);
        fin = true;
        bin = min(it, EO_NITER_BINS - 1);
        nonconv = it >= k.nitermax;
        nonfin = !(isfinite(sig[0]) && isfinite(sig[1]) && isfinite(sig[2]) && isfinite(sig[3]));
        atomicOr(&s_free[lane], 1u << (slot >> 5));  // the slot's contents are dead: no ordering needed
      } else {
        s_it[slot] = it;
        __threadfence_block();  // the slot is complete before another warp's lane of this class can see the bit
        atomicOr(&s_ready[lane], 1u << (slot >> 5));
      }
    }
    // statistics of the points that left the loop: warp-aggregated
    const unsigned fm = __ballot_sync(0xffffffffu, fin);
    if (fm) {
      const unsigned ncm = __ballot_sync(0xffffffffu, nonconv), nfm = __ballot_sync(0xffffffffu, nonfin);
      if (fin) {
        const unsigned grp = __match_any_sync(fm, bin);
        if (lane == __ffs(grp) - 1) atomicAdd(&s_hist[bin], (unsigned)__popc(grp));
      }
      if (lane == 0) {
        atomicAdd(&s_plastic, (unsigned)__popc(fm));
        atomicAdd(&s_inflight, -__popc(fm));
        if (ncm) atomicAdd(&s_nonconv, (unsigned)__popc(ncm));
        if (nfm) atomicAdd(&s_nonfinite, (unsigned)__popc(nfm));
      }
    }
    if (lane == 0) atomicAdd(&s_active, -1);
    __syncwarp();
  }

  // ------------------------------------------------------------------ statistics flush
  __syncthreads();
  for (int b = tid; b < EO_NITER_BINS; b += NWARPS * 32)
    if (s_hist[b]) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->niter_hist[b]), (unsigned long long)s_hist[b]);
  if (tid == 0) {
    if (s_plastic) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_plastic), (unsigned long long)s_plastic);
    if (s_nonconv) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_nonconverged), (unsigned long long)s_nonconv);
    if (s_nonfinite) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_nonfinite), (unsigned long long)s_nonfinite);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    max_res = fmax(max_res, __shfl_xor_sync(0xffffffffu, max_res, o));
    max_it = max(max_it, __shfl_xor_sync(0xffffffffu, max_it, o));
  }
  if (lane == 0) {
    mc_atomic_max_f64(&stats->res_max, max_res);
    mc_atomic_max_f64(&stats->niter_max, (double)max_it);
  }
}

// simple variant: one thread per point, whole Newton loop per thread (divergent).  Kept as the
// baseline the queue scheme is measured against (bench.py --mc-scheme simple) and as a cross-check.
template <bool ASSOC>
__global__ void __launch_bounds__(MC_SIMPLE_THREADS) mc_kernel_simple(const mc_consts k, const mc_ptrs P, const int64_t n) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const eo_d4 e = eo_ld256(P.deps + 4 * i);
  const eo_d4 s = eo_ld256(P.sigma_n + 4 * i);
  const double de[4] = {e.x, e.y, e.z, e.w}, sn[4] = {s.x, s.y, s.z, s.w};
  double Ct[16], sig[4], yl, nr, dl;
  int32_t it;
  mc_point(k, de, sn, Ct, sig, it, yl, nr, dl);
  mc_store_point(P, i, Ct, sig);
  mc_store_aux(P, i, it, yl, nr, dl);
}

static bool g_mc_attr_set[8] = {};

template <bool ASSOC, int AFF, bool LISTED>
static int mc_launch_queue(eo_ctx* ctx, const mc_consts& k, const mc_ptrs& P, int64_t n, size_t smem, unsigned grid) {
  const int a = (ASSOC ? 1 : 0) + 2 * AFF + 4 * (LISTED ? 1 : 0);
  if (!g_mc_attr_set[a]) {
    cudaError_t e = cudaFuncSetAttribute(mc_kernel<ASSOC, AFF, LISTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return eo_fail(ctx, EO_ERR_CUDA, "eo_mc_eval: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    g_mc_attr_set[a] = true;
  }
  int32_t* list = nullptr;
  double* list_yl = nullptr;
  if (LISTED) {
    // pass 1: yield test for every point at full occupancy; plastic points land in the list
    if (n > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_mc_eval: n too large for one launch");
    void* sc = nullptr;
    const size_t yl_off = (size_t(n) * 4 + 255) / 256 * 256;
    int rc = eo_scratch(ctx, yl_off + size_t(n) * 8, &sc);
    if (rc != EO_OK) return rc;
    list = reinterpret_cast<int32_t*>(sc);
    list_yl = reinterpret_cast<double*>(reinterpret_cast<char*>(sc) + yl_off);
    mc_trial_kernel<ASSOC, 0><<<unsigned((n + 255) / 256), 256, 0, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl,
                                                                                   nullptr, mc_tab_src{}, nullptr);
    ctx->launches += 1;
    grid = (unsigned)ctx->sm_count;  // the list length is only known on the device
  }
  mc_kernel<ASSOC, AFF, LISTED><<<grid, MC_THREADS, smem, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl);
  return EO_OK;
}

// default scheme: pass 1 (mc_trial_kernel) + the lane-class Newton kernel over the plastic list
template <bool ASSOC, int NWARPS, int DEPTH>
static int mc_launch_newton(eo_ctx* ctx, const mc_consts& k, const mc_ptrs& P, const int32_t* list, const double* list_yl,
                            const double* list_deps) {
  const size_t smem = size_t(ASSOC ? MC_NF_ASSOC : MC_NF) * (32 * DEPTH) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mc_newton_kernel<ASSOC, NWARPS, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return eo_fail(ctx, EO_ERR_CUDA, "eo_mc_eval: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static const mcn_policy pol = [] {  // EO_MC_POLICY="full,min_active,first_last,sleep_ns" overrides (A/B runs)
    mcn_policy p{MCN_FULL_DEFAULT, MCN_MIN_ACTIVE_DEFAULT, MCN_FIRST_LAST_DEFAULT, 200u};
    if (const char* e = getenv("EO_MC_POLICY")) sscanf(e, "%d,%d,%d,%u", &p.full, &p.min_active, &p.first_last, &p.sleep_ns);
    return p;
  }();
  mc_newton_kernel<ASSOC, NWARPS, DEPTH><<<(unsigned)ctx->sm_count, NWARPS * 32, smem, ctx->s_cmp>>>(k, P, ctx->stats, ctx->work_ctr, list, list_yl, pol,
                                                                                                   list_deps);
  return EO_OK;
}

// default scheme: pass 1 (mc_trial_kernel) + the lane-class Newton kernel over the plastic list.  tab != nullptr: the
// strain increment is tabulated inside pass 1 from the coefficient vector d_u (eo_mc_eval_tabulated).
template <bool ASSOC>
static int mc_launch_classes(eo_ctx* ctx, const mc_consts& k, const mc_ptrs& P, int64_t n, eo_tab* tab = nullptr,
                             const double* d_u = nullptr) {
  // EO_MC_CONFIG selects the CTA shape of pass 2 for A/B runs: 0 = 12 warps x 512 slots, 1 = 16 warps (128 registers) x
  // 512 slots, 2 = 16 warps x 576 slots, 3 = 12 warps x 576 slots (associative flow rule only: 51 fields per slot do not
  // fit otherwise)
  static const int cfg = [] {
    const char* e = getenv("EO_MC_CONFIG");
    return (e && *e >= '0' && *e <= '3') ? *e - '0' : MCN_DEFAULT_CONFIG;
  }();
  if (n > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_mc_eval: n too large for one launch");
  void* sc = nullptr;
  const size_t yl_off = (size_t(n) * 4 + 255) / 256 * 256;
  const size_t de_off = yl_off + (size_t(n) * 8 + 255) / 256 * 256;
  int rc = eo_scratch(ctx, tab ? de_off + size_t(n) * 32 : yl_off + size_t(n) * 8, &sc);
  if (rc != EO_OK) return rc;
  int32_t* list = reinterpret_cast<int32_t*>(sc);
  double* list_yl = reinterpret_cast<double*>(reinterpret_cast<char*>(sc) + yl_off);
  double* list_deps = tab ? reinterpret_cast<double*>(reinterpret_cast<char*>(sc) + de_off) : nullptr;
  const unsigned grid = unsigned((n + 255) / 256);
  if (tab) {
    const mc_tab_src src{tab->dofmap, tab->x_dofmap, tab->x, d_u};
    if (tab->T.nb == 3) mc_trial_kernel<ASSOC, 3><<<grid, 256, 0, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl, tab->d_T, src, list_deps);
    else if (tab->T.nb == 6) mc_trial_kernel<ASSOC, 6><<<grid, 256, 0, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl, tab->d_T, src, list_deps);
    else mc_trial_kernel<ASSOC, 10><<<grid, 256, 0, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl, tab->d_T, src, list_deps);
  } else {
    mc_trial_kernel<ASSOC, 0><<<grid, 256, 0, ctx->s_cmp>>>(k, P, n, ctx->stats, ctx->work_ctr, list, list_yl, nullptr, mc_tab_src{}, nullptr);
  }
  if (cfg == 3 && ASSOC) rc = mc_launch_newton<ASSOC, 12, ASSOC ? 18 : 16>(ctx, k, P, list, list_yl, list_deps);
  else if (cfg == 2 && ASSOC) rc = mc_launch_newton<ASSOC, 16, ASSOC ? 18 : 16>(ctx, k, P, list, list_yl, list_deps);
  else if (cfg >= 1) rc = mc_launch_newton<ASSOC, 16, 16>(ctx, k, P, list, list_yl, list_deps);
  else rc = mc_launch_newton<ASSOC, 12, 16>(ctx, k, P, list, list_yl, list_deps);
  if (rc != EO_OK) return rc;
  ctx->launches += 2;
  return EO_OK;
}

static int mc_launch(eo_ctx* ctx, const mc_consts& k, const mc_ptrs& P, int64_t n, int scheme) {
  if (scheme == 1) {
    const int64_t grid64 = (n + MC_SIMPLE_THREADS - 1) / MC_SIMPLE_THREADS;
    if (grid64 > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_mc_eval: n too large for one launch");
    if (k.assoc)
      mc_kernel_simple<true><<<(unsigned)grid64, MC_SIMPLE_THREADS, 0, ctx->s_cmp>>>(k, P, n);
    else
      mc_kernel_simple<false><<<(unsigned)grid64, MC_SIMPLE_THREADS, 0, ctx->s_cmp>>>(k, P, n);
    ctx->launches += 1;
    return EO_OK;
  }
  const size_t smem = size_t(k.assoc ? MC_NF_ASSOC : MC_NF) * MC_NSLOTS * sizeof(double);
  const int64_t ntiles = (n + MC_TILE - 1) / MC_TILE;
  if (ntiles > 4000000000LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_mc_eval: n too large for one launch");
  const int64_t grid = ntiles < ctx->sm_count ? ntiles : ctx->sm_count;
  cudaError_t e = cudaMemsetAsync(ctx->work_ctr, 0, 256, ctx->s_cmp);
  if (e != cudaSuccess) return eo_fail(ctx, EO_ERR_CUDA, "eo_mc_eval: cudaMemsetAsync: %s", cudaGetErrorString(e));
  if (scheme == 0) return k.assoc ? mc_launch_classes<true>(ctx, k, P, n) : mc_launch_classes<false>(ctx, k, P, n);
  int rc;
  if (scheme == 2)  // one pass, no stage affinity: any warp takes the highest-priority full queue (A/B measurements)
    rc = k.assoc ? mc_launch_queue<true, 0, false>(ctx, k, P, n, smem, (unsigned)grid) : mc_launch_queue<false, 0, false>(ctx, k, P, n, smem, (unsigned)grid);
  else if (scheme == 3)  // one pass with stage affinity: the yield test is the scheduler's T stage (A/B measurements)
    rc = k.assoc ? mc_launch_queue<true, 1, false>(ctx, k, P, n, smem, (unsigned)grid) : mc_launch_queue<false, 1, false>(ctx, k, P, n, smem, (unsigned)grid);
  else  // scheme 4: two passes with the ring scheduler of round 1 (A/B measurements)
    rc = k.assoc ? mc_launch_queue<true, 1, true>(ctx, k, P, n, smem, (unsigned)grid) : mc_launch_queue<false, 1, true>(ctx, k, P, n, smem, (unsigned)grid);
  if (rc != EO_OK) return rc;
  ctx->launches += 1;
  return EO_OK;
}

extern "C" {

int eo_mc_eval(eo_ctx* ctx, const eo_mc_params* prm, const double* deps, const double* sigma_n, double* C_tang,
               double* sigma, int32_t* niter, double* yielding, double* norm_res, double* dlambda, int64_t n) {
  static const int def_scheme = [] {  // EO_MC_SCHEME overrides the default execution scheme (A/B runs of the test-suite)
    const char* e = getenv("EO_MC_SCHEME");
    return (e && *e >= '0' && *e <= '4') ? *e - '0' : 0;
  }();
  return eo_mc_eval_scheme(ctx, prm, deps, sigma_n, C_tang, sigma, niter, yielding, norm_res, dlambda, n, def_scheme);
}

int eo_mc_eval_scheme(eo_ctx* ctx, const eo_mc_params* prm, const double* deps, const double* sigma_n, double* C_tang,
                      double* sigma, int32_t* niter, double* yielding, double* norm_res, double* dlambda, int64_t n,
                      int scheme) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_mc_eval: ctx is NULL");
  EO_REQUIRE(ctx, prm != nullptr, "eo_mc_eval: prm is NULL");
  EO_REQUIRE(ctx, n >= 0, "eo_mc_eval: n < 0");
  EO_REQUIRE(ctx, scheme >= 0 && scheme <= 4, "eo_mc_eval: unknown scheme");
  EO_REQUIRE(ctx, prm->Nitermax >= 0 && prm->Nitermax <= 200,
             "eo_mc_eval: Nitermax must be in [0, 200] (histogram bins of eo_stats)");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, deps && sigma_n && C_tang && sigma, "eo_mc_eval: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  mc_params_in pin{prm->E, prm->nu, prm->c, prm->phi, prm->psi, prm->theta_T, prm->a, prm->tol, prm->Nitermax};
  mc_consts k;
  mc_make_consts(pin, k);
  eo_arg args[8] = {{deps, 32, false},  {sigma_n, 32, false}, {C_tang, 128, true},  {sigma, 32, true},
                    {niter, 4, true},   {yielding, 8, true},  {norm_res, 8, true}, {dlambda, 8, true}};
  return eo_run_streamed(ctx, args, 8, n, [&](void** a, int64_t m, int64_t) {
    for (int i = 0; i < 4; ++i)
      if (!eo_aligned(a[i], 32)) return eo_fail(ctx, EO_ERR_INVALID, "eo_mc_eval: deps/sigma_n/C_tang/sigma must be 32-byte aligned");
    mc_ptrs P{(const double*)a[0], (const double*)a[1], (double*)a[2], (double*)a[3],
              (int32_t*)a[4],      (double*)a[5],       (double*)a[6], (double*)a[7]};
    return mc_launch(ctx, k, P, m, scheme);
  });
}

int eo_mc_eval_tabulated(eo_ctx* ctx, const eo_mc_params* prm, eo_tab* tab, const double* u, const double* sigma_n,
                         double* C_tang, double* sigma, int32_t* niter, double* yielding, double* norm_res, double* dlambda) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_mc_eval_tabulated: ctx is NULL");
  EO_REQUIRE(ctx, prm && tab, "eo_mc_eval_tabulated: NULL argument");
  EO_REQUIRE(ctx, tab->ctx == ctx, "eo_mc_eval_tabulated: the tabulation handle lives on another context");
  EO_REQUIRE(ctx, tab->T.gdim == 2 && tab->T.bs == 2 && (tab->T.nb == 3 || tab->T.nb == 6 || tab->T.nb == 10),
             "eo_mc_eval_tabulated: needs a P1/P2/P3 vector field on triangles (plane-strain Mandel strain)");
  EO_REQUIRE(ctx, prm->Nitermax >= 0 && prm->Nitermax <= 200, "eo_mc_eval_tabulated: Nitermax must be in [0, 200]");
  const int64_t n = tab->n_cells * tab->T.nq;
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, u && sigma_n && C_tang && sigma, "eo_mc_eval_tabulated: NULL array");
  EO_REQUIRE(ctx, eo_is_device_ptr(sigma_n) && eo_is_device_ptr(C_tang) && eo_is_device_ptr(sigma) &&
                      (!niter || eo_is_device_ptr(niter)) && (!yielding || eo_is_device_ptr(yielding)) &&
                      (!norm_res || eo_is_device_ptr(norm_res)) && (!dlambda || eo_is_device_ptr(dlambda)),
             "eo_mc_eval_tabulated: history and outputs must be device memory (the coefficient vector may be host memory)");
  EO_REQUIRE(ctx, eo_aligned(sigma_n, 32) && eo_aligned(C_tang, 32) && eo_aligned(sigma, 32),
             "eo_mc_eval_tabulated: arrays must be 32-byte aligned");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const double* d_u = nullptr;
  int rc = eo_tab_stage_u(tab, u, &d_u);
  if (rc != EO_OK) return rc;
  mc_params_in pin{prm->E, prm->nu, prm->c, prm->phi, prm->psi, prm->theta_T, prm->a, prm->tol, prm->Nitermax};
  mc_consts k;
  mc_make_consts(pin, k);
  EO_CUDA(ctx, cudaMemsetAsync(ctx->work_ctr, 0, 256, ctx->s_cmp));
  const mc_ptrs P{nullptr, sigma_n, C_tang, sigma, niter, yielding, norm_res, dlambda};
  rc = k.assoc ? mc_launch_classes<true>(ctx, k, P, n, tab, d_u) : mc_launch_classes<false>(ctx, k, P, n, tab, d_u);
  if (rc != EO_OK) return rc;
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

}  // extern "C"
