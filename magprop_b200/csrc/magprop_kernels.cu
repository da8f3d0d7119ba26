// magprop_kernels.cu -- sm_100a kernels and the C ABI of include/magprop_b200.h.
//
// FP64 throughout, no tensor cores -- the path is a scalar ODE solve per walker, not a contraction.
// One likelihood evaluation runs as a short pipeline of launches on one stream:
//   setup_kernel      prior -> parameters -> per-walker constants -> initial step      one thread per walker
//   advance_kernel    the spin integration (Dormand-Prince 5(4), dense output at the   persistent warps; every lane
//                     nodes the data need); stiff walkers are queued for ...            pulls its next walker off a
//   advance_kernel<STIFF>  ... the implicit (Radau IIA) integrator                      queue when its own is done
//   reduce_*_kernel   luminosity -> interpolation -> chi-square -> lnprob (or model /  thread per walker, or a warp
//                     light-curve output; or the stretch move's accept step)           per walker with a shuffle-
//                                                                                       reduced chi-square
// plus rhs_kernel (the coupled reference RHS for ODEs()/odes() callers), the ensemble-order and peer-barrier
// helpers of the device-resident sampler, and dfma_peak_kernel (FP64 FMA roofline denominator).
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "magprop_host.hpp"
#include "magprop_rng.cuh"

namespace mp {

constexpr unsigned kFull = 0xffffffffu;

// Warp-aggregated append: the lanes with `want` get consecutive slots of a list whose length is *count.
__device__ __forceinline__ int warp_append(bool want, int* count) {
  const unsigned m = __ballot_sync(kFull, want);
  if (!m) return -1;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(kFull, base, leader);
  return want ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}


// What every stage of a launch needs to know (passed by value in the kernel parameters).
struct Problem {
  Spec sp;
  DataView dv;          // nodes (+ the dataset for lnprob / model-at-data)
  const int* dat_orig;  // sorted datum -> caller's index (kModeModelAtData)
  double lower[MP_MAX_NDIM], upper[MP_MAX_NDIM];
  int prior_enabled;
  int ndim;
};

// Device work space of one launch (a slab of walkers), owned by the handle's lane.  Layouts follow the readers:
//   recs   the stage-1 records as a structure of arrays -- 8-byte word k of walker i at recs[k*stride + i] -- so
//          that the one-thread-per-walker kernels read and write them coalesced
//   ybuf   node j of walker i at ybuf[i*ws + j*ns]: walker-minor (ws = 1, ns = stride) when stage 3 runs one
//          thread per walker (small datasets), node-minor (ws = Nn, ns = 1) when it runs one warp per walker
struct Work {
  int W;               // walkers in the slab
  int stride;          // slab capacity
  size_t ws, ns;
  unsigned long long* recs;
  double* ybuf;        // stage 2 -> stage 3: y = omega^-2 at the nodes
  int* status;         // [W]
  int* n_rhs;          // [W]
  int* work;           // [W] walkers for the explicit integrator
  StiffRec* squeue;    // [W] walkers for the implicit integrator
  int* counters;       // [0] walkers in `work`, [1] next to hand out, [2] walkers in `squeue`, [3] next to hand out,
                       // [4..7] extent of the ensemble in the ordering's bins: max(im + 1), max(256 - im), max(ie + 1), max(32 - ie),
                       // [8..11] the extent the host has been told
  int track;           // 1: the setup kernel measures [4..7] (launches large enough to be ordered)
  int* hint;           // (track) where the host reads [4..7] later: mapped pinned memory, written by the implicit launch's first thread
  // ordering of `work` (large launches): per-walker bucket key and the bucket histogram / cursors
  int* key;            // [W] bucket key of every walker
  int* hist;           // [kOrderBuckets + 3]: bucket counts / offsets, "tight" flag, tight cursor, block ticket
  int* slot_wid;       // [W] or null: the stages index their work space by SLOT; slot s holds walker slot_wid[s]
};                     //          (null: slot = walker, launches too small to be worth ordering)

// The order in which the integrator takes the walkers of a large launch.  Lanes that start together run through
// the same phases of the integration together only if their walkers are alike, so the walkers of a large launch are
// given SLOTS -- the index under which all three stages keep a walker's record, node values and status, so that
// neighbouring lanes also touch neighbouring memory -- in the order of a key of
// the two parameter combinations that set a walker's timeline: the mass that flows through the disc, M_disc * delta
// (1/24-decade bins, DESCENDING: the number of steps grows with it -- correlation 0.83..0.94 with log(steps) on
// posterior-like ensembles of the four synthetic datasets -- and a launch whose long integrations start last ends
// with a tail as long as one of them, `tools/sched_sim.py`) and, within a bin, the fallback time scale epsilon
// (quarter-decade bins).  A counting sort: histogram in order_key_kernel, one scan, one scatter.
// Measured on 2^18 walkers against the unordered list: prior-uniform +13 %, posterior-like spreads +10..30 %,
// a 1e-4 ball unchanged; results do not depend on the order (tested).
constexpr int kOrderBuckets = 32 * 256;
#ifndef MP_ORDER_MIN_WALKERS
#define MP_ORDER_MIN_WALKERS 8192
#endif
constexpr int kOrderMinWalkers = MP_ORDER_MIN_WALKERS;
constexpr int kCoopSmallWalkers = 2048;   // launches up to this size run stage 3 with a warp per walker (see reduce_coop_kernel)
constexpr int kOrderTightBuckets = 8;
__device__ __forceinline__ void order_bins(const Spec& sp, const double* th, int& ie, int& im) {
  const float le = ((sp.unlog_mask >> 4) & 1) ? (float)th[4] : __log10f((float)th[4]);
  const float lm = ((sp.unlog_mask >> 2) & 1) ? (float)th[2] : __log10f((float)th[2]);
  const float ld = ((sp.unlog_mask >> 5) & 1) ? (float)th[5] : __log10f((float)th[5]);
  const float eb = fminf(fmaxf((le + 4.0f) * 4.0f, 0.0f), 31.0f);
  const float mb = fminf(fmaxf((lm + ld + 10.0f) * 24.0f, 0.0f), 255.0f);
  ie = (eb == eb) ? (int)eb : 0;
  im = (mb == mb) ? (int)mb : 0;
}
__device__ __forceinline__ int order_key_of(const Spec& sp, const double* th) {
  int ie, im;
  order_bins(sp, th, ie, im);
  // descending in M*delta (the long integrations first), zig-zag in epsilon: neighbouring buckets across an M*delta boundary are alike
  return (255 - im) * 32 + ((im & 1) ? 31 - ie : ie);
}
// An ensemble inside a few neighbouring buckets gains nothing from the ordering (launch_eval skips it).
constexpr int kTightMdBins = 2, kTightEpsBins = 1;
// The threads of a block claim slots of their keys' counters: lanes with the same key share one atomic, and when
// the whole block holds one key (a tight ensemble: every walker in the same bucket) so do its warps -- otherwise
// thousands of warps would queue on one address.  Every thread of the block must call (it synchronises).
__device__ __forceinline__ int block_claim_by_key(bool want, int key, int* counters) {
  __shared__ int s_key[32], s_cnt[32], s_base;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const unsigned wanting = __ballot_sync(kFull, want);
  const unsigned peers = __match_any_sync(kFull, want ? key : -1);
  // this warp's key, from its first wanting lane: -1 nothing to claim, -2 mixed keys
  const int first = wanting ? __ffs(wanting) - 1 : 0;
  const int fkey = __shfl_sync(kFull, key, first);
  const unsigned fpeers = __shfl_sync(kFull, peers, first);
  if (lane == 0) {
    s_cnt[wrp] = __popc(wanting);
    s_key[wrp] = !wanting ? -1 : ((fpeers == wanting) ? fkey : -2);
  }
  __syncthreads();
  int common = -1, before = 0, total = 0;
  bool uniform = true;
  for (int w = 0; w < nw; ++w) {
    const int kw = s_key[w];
    if (kw == -1) continue;
    if (kw == -2 || (common >= 0 && kw != common)) uniform = false;
    if (common < 0) common = kw;
    if (w < wrp) before += s_cnt[w];
    total += s_cnt[w];
  }
  int pos = -1;
  if (uniform && common >= 0) {
    if (threadIdx.x == 0) s_base = atomicAdd(counters + common, total);
    __syncthreads();
    if (want) pos = s_base + before + __popc(wanting & ((1u << lane) - 1u));
  } else {
    __syncthreads();
    if (want) {
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(counters + key, __popc(peers));
      base = __shfl_sync(peers, base, leader);
      pos = base + __popc(peers & ((1u << lane) - 1u));
    }
  }
  return pos;
}
// Exclusive scan of the bucket histogram in place (hist[b] becomes the first slot of bucket b), by the NT threads of
// one block.  An ensemble that occupies only a handful of buckets is tight enough to be taken as it comes:
// hist[kOrderBuckets] tells the scatter so, which then hands out the slots in index order (cursor: hist[kOrderBuckets + 1]).
template <int NT>
__device__ __forceinline__ void scan_buckets(int* __restrict__ hist) {
  __shared__ int part[NT];
  __shared__ int occupied;
  constexpr int per = kOrderBuckets / NT;
  const int base = threadIdx.x * per;
  int sum = 0, nz = 0;
  if (threadIdx.x == 0) occupied = 0;
  for (int c = 0; c < per; ++c) { const int v = __ldcg(hist + base + c); sum += v; nz += v != 0; }   // (written by other blocks' atomics: read at L2)
  part[threadIdx.x] = sum;
  __syncthreads();
  if (nz) atomicAdd(&occupied, nz);
  __syncthreads();
  if (occupied <= kOrderTightBuckets) {
    if (threadIdx.x == 0) hist[kOrderBuckets] = 1;
    return;
  }
  for (int off = 1; off < NT; off <<= 1) {
    const int v = (threadIdx.x >= off) ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;
  for (int c = 0; c < per; ++c) { const int v = __ldcg(hist + base + c); hist[base + c] = run; run += v; }
}
__global__ void order_scatter_kernel(int W, const int* __restrict__ key, int* __restrict__ cursor, int* __restrict__ slot_wid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int kk = (i < W) ? key[i] : -1;
  if (cursor[kOrderBuckets]) {                       // tight ensemble: as it comes
    const int we = warp_append(kk >= 0, cursor + kOrderBuckets + 1);
    if (kk >= 0) slot_wid[we] = i;
    return;
  }
  const int pos = block_claim_by_key(kk >= 0, kk, cursor);
  if (kk >= 0) slot_wid[pos] = i;
}

// The stretch move wrapped around an evaluation (mp_stretch_half_step / mp_ensemble_half_step):
//   z = ((a-1) u + 1)^2 / a ; q = c - (c - s) z ; accept iff (ndim-1) ln z + lp(q) - lp(s) > ln u'
// (Goodman & Weare 2010; emcee's RedBlueMove).  setup_kernel forms the proposals, the reduce kernels accept.
struct Move {
  double* coords;        // [nwalkers][ndim], updated in place
  double* lnp;           // [nwalkers]
  // who moves against whom: explicit index lists (mp_stretch_half_step) ...
  const int* active;     // [n_active] walkers to move (disjoint from complement), or null
  const int* complement; // [n_complement]
  // ... or positions of the ensemble order (mp_ensemble_half_step): mover i is walker P(pos0 + i), its
  // partner is drawn from P(cpos0 + [0, n_complement))
  SplitPerm perm;
  int pos0, cpos0;
  int n_complement;
  double a;
  uint64_t seed, step;   // step: the half-step counter (RNG counter word)
  double* prop;          // [n_active][ndim] the proposals (work space)
  int* accepted;         // [nwalkers] counters (may be null)
  int* status;           // [nwalkers] MP_WALKER_* bits of the latest proposal (may be null)
  int* n_rhs;            // [nwalkers] (may be null)
  // Sharded ensembles with peer-mapped replicas (NVLink): a rank writes the rows it moves into its OWN replica
  // only, and reads a row it needs -- its movers' current positions, their partners -- from the replica of the
  // rank that moved that walker last, which every rank can work out (the ensemble order is a keyed permutation):
  //   moved earlier in this step (the partners of half-step 1)  ->  rank (position within the half) / n_mine
  //   otherwise                                                 ->  the same rule with LAST step's order, unless the
  //                                                                 replicas were synchronised since (`synced`)
  // No collective, no remote stores; one flag barrier per half-step (mp_peer_barrier) orders the reads after the writes.
  int n_peers;
  const double* peer_coords[MP_MAX_PEERS];   // the other ranks' replicas, in rank order without this rank
  const double* peer_lnp[MP_MAX_PEERS];
  SplitPerm prev_perm;   // last step's ensemble order
  int rank, n_mine, half, split, synced;
  double* pack_out;      // [n_active][ndim+1]: every mover's (row, lnp) after the move, or null
  double* bad_rows;      // proposals whose likelihood was not finite ({GRB}_bad.csv, mcmc_eqns.py:72-79)
  int* bad_count;
  int bad_capacity;
};

struct Draw {
  double z, ua;
  int me, partner, pj;
};
// Which rank's replica holds walker w's current row (see Move).
__device__ __forceinline__ int home_before_this_step(const Move& m, int w) {
  if (m.synced) return m.rank;
  return (int)(perm_inv(m.prev_perm, (uint32_t)w) % (uint32_t)m.half) / m.n_mine;
}
__device__ __forceinline__ const double* replica_coords(const Move& m, int home) {
  return home == m.rank ? m.coords : m.peer_coords[home < m.rank ? home : home - 1];
}
__device__ __forceinline__ const double* replica_lnp(const Move& m, int home) {
  return home == m.rank ? m.lnp : m.peer_lnp[home < m.rank ? home : home - 1];
}
// counter = (half-step, walker); two Philox blocks give u_z, u_partner, u_accept
__device__ __forceinline__ Draw draw_for(const Move& m, int i) {
  Draw d;
  d.me = m.active ? m.active[i] : (int)perm_at(m.perm, (uint32_t)(m.pos0 + i));
  const Philox r0 = philox4x32_10((uint32_t)m.step, (uint32_t)(m.step >> 32), (uint32_t)d.me, 0u,
                                  (uint32_t)m.seed, (uint32_t)(m.seed >> 32));
  const Philox r1 = philox4x32_10((uint32_t)m.step, (uint32_t)(m.step >> 32), (uint32_t)d.me, 1u,
                                  (uint32_t)m.seed, (uint32_t)(m.seed >> 32));
  const double uz = u01(r0.c[0], r0.c[1]);
  const double up = u01(r0.c[2], r0.c[3]);
  d.ua = u01(r1.c[0], r1.c[1]);
  const double zr = __dadd_rn(__dmul_rn(m.a - 1.0, uz), 1.0);
  d.z = __ddiv_rn(__dmul_rn(zr, zr), m.a);
  int pj = (int)(up * m.n_complement);
  if (pj >= m.n_complement) pj = m.n_complement - 1;
  d.partner = m.complement ? m.complement[pj] : (int)perm_at(m.perm, (uint32_t)(m.cpos0 + pj));
  d.pj = pj;
  return d;
}

constexpr int kRecWords = (int)(sizeof(WalkerRec) / 8);
static_assert(sizeof(WalkerRec) % 8 == 0, "WalkerRec is moved as 8-byte words");
__device__ __forceinline__ void rec_store(const Work& k, int i, const WalkerRec& r) {
  unsigned long long tmp[kRecWords];
  memcpy(tmp, &r, sizeof(r));
#pragma unroll
  for (int c = 0; c < kRecWords; ++c) k.recs[(size_t)c * k.stride + i] = tmp[c];
}
__device__ __forceinline__ void rec_load(const Work& k, int i, WalkerRec& r) {
  unsigned long long tmp[kRecWords];
#pragma unroll
  for (int c = 0; c < kRecWords; ++c) tmp[c] = k.recs[(size_t)c * k.stride + i];
  memcpy(&r, tmp, sizeof(r));
}

// Only what a stage reads of a record: stage 3 needs disc_mass and the luminosity stage (19 of the 55 words),
// the integrators the disc-mass and torque constants (16) -- a lane taking a walker loads them one lane at a time.
#define MP_WFIELD(name) \
  w.name = __longlong_as_double((long long)k.recs[((offsetof(WalkerRec, w) + offsetof(Walker, name)) / 8) * (size_t)k.stride + i])
#define MP_RWORD(member) k.recs[(offsetof(WalkerRec, member) / 8) * (size_t)k.stride + i]
__device__ __forceinline__ void rec_load_step(const Work& k, int i, Walker& w) {
  MP_WFIELD(inv_tv); MP_WFIELD(eps); MP_WFIELD(u0); MP_WFIELD(K); MP_WFIELD(C); MP_WFIELD(u_late); MP_WFIELD(Kq);
  MP_WFIELD(sqrtA); MP_WFIELD(KqA); MP_WFIELD(tvI); MP_WFIELD(g_sqrtA); MP_WFIELD(g_tvI); MP_WFIELD(Cdip_I); MP_WFIELD(Cdip_I2);
  MP_WFIELD(KtvI); MP_WFIELD(M_init);
  w.bad = 0;
}
__device__ __forceinline__ void rec_load_lum(const Work& k, int i, Walker& w) {
  MP_WFIELD(inv_tv); MP_WFIELD(eps); MP_WFIELD(u0); MP_WFIELD(K); MP_WFIELD(C); MP_WFIELD(M_init); MP_WFIELD(u_late);
  MP_WFIELD(l_inv_tv); MP_WFIELD(l_Ccap); MP_WFIELD(l_kc); MP_WFIELD(l_sqrtA); MP_WFIELD(l_sGMkc); MP_WFIELD(l_GM_kc);
  MP_WFIELD(Ldip_coef); MP_WFIELD(dipeff); MP_WFIELD(propeff); MP_WFIELD(f_beam); MP_WFIELD(omega0);
  w.bad = (int)k.recs[((offsetof(WalkerRec, w) + offsetof(Walker, bad)) / 8) * (size_t)k.stride + i];
}

// ---- stage 1: setup ---------------------------------------------------------------------------------
// The stretch proposal of mover i (see Move): q = c - (c - s) z, kept in m.prop[i] for the accept step.  With
// peer-mapped replicas the mover's current row is first brought into the local replica (it is about to be moved
// here) and the partner's row is read from its home.
__device__ __forceinline__ void propose(const Move& m, int i, int ndim, double* th) {
  const Draw d = draw_for(m, i);
  const double* pc = m.coords;
  if (m.n_peers > 0) {
    const int hm = home_before_this_step(m, d.me);
    if (hm != m.rank) {
      const double* src = replica_coords(m, hm);
      for (int c = 0; c < ndim; ++c) m.coords[(size_t)d.me * ndim + c] = src[(size_t)d.me * ndim + c];
      m.lnp[d.me] = replica_lnp(m, hm)[d.me];
    }
    pc = replica_coords(m, m.split == 1 ? d.pj / m.n_mine : home_before_this_step(m, d.partner));
  }
  for (int c = 0; c < ndim; ++c) {
    const double cc = pc[(size_t)d.partner * ndim + c];
    const double x = m.coords[(size_t)d.me * ndim + c];
    th[c] = __dadd_rn(cc, -__dmul_rn(__dadd_rn(cc, -x), d.z));   // no FMA contraction: reproducible on the host
    m.prop[(size_t)i * ndim + c] = th[c];
  }
}

// Large launches, before the setup: every walker's bucket key and the bucket histogram (MOVE: the proposals are
// formed here, the setup then reads them from m.prop).
template <bool MOVE>
__global__ void __launch_bounds__(128) order_key_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k,
                                                        const double* __restrict__ theta, const __grid_constant__ Move m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool have = i < k.W;
  double th[MP_MAX_NDIM];
  int kk = -1;
  if (have) {
    if (MOVE) propose(m, i, p.ndim, th);
    else
      for (int c = 0; c < p.ndim; ++c) th[c] = theta[(size_t)i * p.ndim + c];
    kk = order_key_of(p.sp, th);
    k.key[i] = kk;
  }
  block_claim_by_key(have, kk, k.hist);
  // the last block to get here turns the histogram into bucket offsets (saves a launch)
  __shared__ int last;
  __threadfence();
  if (threadIdx.x == 0) last = atomicAdd(k.hist + kOrderBuckets + 2, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    scan_buckets<128>(k.hist);
  }
}

// MOVE = false: walker i's parameters are theta[i][:].  MOVE = true: walker i is mover i of a stretch-move
// half-step and its parameters are the proposal q.  One thread per SLOT (= walker unless the launch is ordered).
template <bool MOVE>
__global__ void __launch_bounds__(128, 4) setup_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k,
                                                    const double* __restrict__ theta, const __grid_constant__ Move m) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  const bool have = sl < k.W;
  const int ndim = p.ndim;
  double th[MP_MAX_NDIM];
  if (have) {
    const int i = k.slot_wid ? k.slot_wid[sl] : sl;
    if (MOVE && !k.slot_wid) {
      propose(m, i, ndim, th);
    } else {
      const double* src = MOVE ? m.prop : theta;
      for (int c = 0; c < ndim; ++c) th[c] = src[(size_t)i * ndim + c];
    }
  }
  bool to_explicit = false, to_implicit = false;
  double y0 = 0.0, h0 = 0.0;
  int n_rhs0 = 0;
  if (have) {
    WalkerRec r;
    const double t_end = p.dv.n_nodes > 0 ? p.dv.node_t[p.dv.n_nodes - 1] : p.dv.t_start;
    prepare_walker(p.sp, th, ndim, p.prior_enabled != 0, p.lower, p.upper, p.dv.t_start, t_end, r);
    if (!(r.status & kWalkerPriorReject)) rec_store(k, sl, r);
    y0 = r.y0; h0 = r.h0; n_rhs0 = r.n_rhs;
    k.status[sl] = r.status;
    k.n_rhs[sl] = r.n_rhs;
    // (a walker whose initialisation failed still goes to the integrator, which marks its nodes)
    const bool go = (r.status == kWalkerOk || r.status == kWalkerIntegratorFail) && p.dv.n_nodes > 0;
    to_implicit = go && p.sp.bucciantini && r.status == kWalkerOk;
    to_explicit = go && !to_implicit;
  }
  if (k.track && (blockIdx.x & 15) == 0) {
    // how far the ensemble reaches in the ordering's bins (the host decides from it whether the NEXT launch is
    // ordered); every 16th block looks -- a sample of >= 512 walkers is ample for the purpose
    const bool in = to_explicit || to_implicit;
    int ie = 0, im = 0;
    if (in) order_bins(p.sp, th, ie, im);
    const int v[4] = {in ? im + 1 : 0, in ? 256 - im : 0, in ? ie + 1 : 0, in ? 32 - ie : 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = __reduce_max_sync(kFull, v[q]);
      if ((threadIdx.x & 31) == 0 && m > __ldcg(k.counters + 4 + q)) atomicMax(k.counters + 4 + q, m);
    }
  }
  const int we = warp_append(to_explicit, k.counters + 0);      // (threads run in slot order: so does the work list)
  if (to_explicit) k.work[we] = sl;
  const int wi = warp_append(to_implicit, k.counters + 2);
  if (to_implicit) {
    StiffRec q;
    q.t = p.dv.t_start; q.y = y0; q.h = h0; q.wid = sl; q.jn = 0; q.n_rhs = n_rhs0; q.n_steps = 0;
    k.squeue[wi] = q;
  }
}

// ---- stage 2: advance -------------------------------------------------------------------------------
// Persistent warps.  Every trip of the loop is ONE integrator step for every lane that holds a walker; a lane
// whose walker is finished (all nodes delivered), failed or -- explicit variant -- turned stiff hands it on
// and takes the next walker off the queue at the top of the next trip, so the lanes of a warp stay busy
// whatever the spread of step counts in the ensemble (40 .. 10^3 over the prior box).  The cheap, divergent
// parts (delivering nodes, hand-over, taking a walker) sit between the steps.
// Resident blocks per SM the explicit integrator is compiled for: 10 x 64 threads = 20 warps at 96 registers.
// (With the luminosity stage out of this kernel the spills at 96 registers are 86 bytes; measured against 8 blocks /
// 128 registers: +7 % on 2^18-walker launches, +5 % through the host-pointer path; 12 and 14 blocks no better.)
#ifndef MP_MIN_BLOCKS_32
#define MP_MIN_BLOCKS_32 20
#endif
#ifndef MP_MIN_BLOCKS_64
#define MP_MIN_BLOCKS_64 10
#endif
// Resident warps per SM the implicit variant is compiled for: 12 (166 registers, no spills), and 16 (128 registers)
// for launches large enough that its queue is longer than one wave of the 12-warp build -- 75 904 stiff walkers of
// 2^18 prior-uniform ones on 56 832 lanes are 1.34 waves, on 75 776 lanes one: 7.13 -> 6.71 ms per launch, 10^6
// walkers 22.2 -> 21.5; a single wave (2^16 walkers) is faster on the 12-warp build, 4.78 against 5.19 ms.
#ifndef MP_STIFF_MIN_WARPS
#define MP_STIFF_MIN_WARPS 12
#endif
constexpr int kStiffWarpsWide = 16;
constexpr int kStiffWideMinWalkers = 196608;
constexpr int kNodeSmemDoubles = 512;      // node times staged per block when they fit (4 KB)

template <bool STIFF, int BLOCK, int STIFF_WARPS = MP_STIFF_MIN_WARPS>
__global__ void __launch_bounds__(BLOCK, STIFF ? (STIFF_WARPS * 32 / BLOCK) : (BLOCK == 32 ? MP_MIN_BLOCKS_32 : MP_MIN_BLOCKS_64))
advance_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k) {
  __shared__ double s_nodes[kNodeSmemDoubles];
  const int Nn = p.dv.n_nodes;
  const double* node_t = p.dv.node_t;
  if (Nn <= kNodeSmemDoubles) {
    for (int j = threadIdx.x; j < Nn; j += BLOCK) s_nodes[j] = node_t[j];
    node_t = s_nodes;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  if (STIFF && k.track && blockIdx.x == 0 && threadIdx.x == 0)        // (see launch_eval: the next launch's hint;
    for (int q = 0; q < 4; ++q) {                                     // written over PCIe only when it changes)
      const int v = k.counters[4 + q];
      if (v != k.counters[8 + q]) { k.counters[8 + q] = v; k.hint[q] = v; }
    }
  const int n_items = k.counters[STIFF ? 2 : 0];
  int* next = k.counters + (STIFF ? 3 : 1);
  const double t_end = ldd(node_t + (Nn - 1));
  Walker w;
  Integrator in;
  in.status = kWalkerOk; in.stiff = 0; in.n_rhs = 0; in.t = p.dv.t_start;
  int wid = 0, jn = 0;
  double* row = k.ybuf;
  bool have = false, empty = false;
  int waited = 0;
  for (;;) {
    // ---- take walkers
    // A lane that is done waits a few trips for the other lanes of its warp before it takes a new walker:
    // lanes that start together run through the same phases of the integration together (capped / uncapped
    // Alfven radius, saturated / unsaturated tanh, kinks on the same trips), which is worth more than the idle
    // trips when the ensemble is tight (+10 % at a posterior-like spread of 0.01..0.05, measured), and when it
    // is not -- prior-uniform, where the step counts differ tenfold -- the wait is over after MP_REFILL_PATIENCE trips.
    const unsigned idle = __ballot_sync(kFull, !have);
#ifndef MP_REFILL_PATIENCE
#define MP_REFILL_PATIENCE 32
#endif
    waited = idle ? waited + 1 : 0;
    // (one warp-wide vote per trip: whether any lane holds a walker follows from `idle` unless lanes were just refilled)
    if (!(idle && !empty && (idle == kFull || waited > MP_REFILL_PATIENCE))) {
      if (idle == kFull) break;
    } else {
      waited = 0;
      const int leader = __ffs(idle) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(next, __popc(idle));
      base = __shfl_sync(kFull, base, leader);
      if (base + __popc(idle) >= n_items) empty = true;          // (uniform across the warp)
      const int my = base + __popc(idle & ((1u << lane) - 1u));
      if (!have && my < n_items) {
        have = true;
        if (STIFF) {
          const StiffRec q = k.squeue[my];
          wid = q.wid;
          rec_load_step(k, wid, w);
          integrator_load_stiff(q, in);
          jn = q.jn;
        } else {
          wid = k.work[my];
          rec_load_step(k, wid, w);
          const int i = wid;
          WalkerRec r;                                             // (its Walker part stays unset: see rec_load_step)
          r.y0 = __longlong_as_double((long long)MP_RWORD(y0));
          r.h0 = __longlong_as_double((long long)MP_RWORD(h0));
          r.k1 = __longlong_as_double((long long)MP_RWORD(k1));
          const unsigned long long rs = MP_RWORD(regime0);         // regime0 | status << 32
          r.regime0 = (unsigned)rs;
          r.status = (int)(rs >> 32);
          r.n_rhs = (int)(unsigned)MP_RWORD(n_rhs);
          integrator_load(r, w.C, p.dv.t_start, in);
          if (r.status != kWalkerOk) in.status = kWalkerIntegratorFail;
          jn = 0;
        }
        row = k.ybuf + (size_t)wid * k.ws;
        jn = drain_nodes<STIFF>(in, jn, Nn, node_t, row, k.ns);      // a node at the starting time
      }
      if (!__any_sync(kFull, have)) break;
    }
    // ---- one step for every lane that holds a walker (a fresh walker whose nodes are all delivered
    // already, or whose initialisation failed, skips it)
    if (have && jn < Nn && in.status == kWalkerOk) {
      if (STIFF) radau_step(p.sp, w, t_end, in);
      else integrator_step(p.sp, w, t_end, in);
      jn = drain_nodes<STIFF>(in, jn, Nn, node_t, row, k.ns);
    }
    // ---- hand finished walkers on
    if (have) {
      const bool failed = in.status != kWalkerOk;
      const bool stiff = !STIFF && in.stiff && !failed && jn < Nn;
      if (jn >= Nn || failed || stiff) {
        if (failed)
          for (; jn < Nn; ++jn) row[jn * k.ns] = NAN;            // stage 3 sees no solution there
        if (stiff) {
          // (explicit variant) the implicit integrator picks the walker up where this one stopped
          StiffRec q;
          q.t = in.t; q.y = in.omega; q.h = in.h; q.wid = wid; q.jn = jn; q.n_rhs = in.n_rhs; q.n_steps = in.n_steps;
          k.squeue[atomicAdd(k.counters + 2, 1)] = q;
        } else {
          k.status[wid] |= in.status;
          k.n_rhs[wid] = in.n_rhs;
        }
        have = false;
      }
    }
  }
}

// ---- stage 3: reduce ----------------------------------------------------------------------------------
// The dataset a block works against -- node times and, per datum, y/yerr, 1e-50/yerr, x - t_lo, the
// interpolation weight and the lower-node index -- staged in shared memory when it fits the budget below
// (the synthetic datasets take 2.2 KB).
constexpr int kDataSmemDoubles = 768;      // 6 KB per block
__device__ __forceinline__ bool stage_data(const DataView& g, DataView& s, double* buf, int nthreads) {
  const int Nn = g.n_nodes, D = g.n_data;
  const int need = Nn + 4 * D + (D + 1) / 2;
  if (need > kDataSmemDoubles) return false;
  double* q = buf;
  double* nt = q; q += Nn;
  double* ys = q; q += D;
  double* c = q; q += D;
  double* dx = q; q += D;
  double* w = q; q += D;
  int* lo = reinterpret_cast<int*>(q);
  for (int i = threadIdx.x; i < Nn; i += nthreads) nt[i] = g.node_t[i];
  for (int i = threadIdx.x; i < D; i += nthreads) {
    ys[i] = g.dat_ys[i]; c[i] = g.dat_c[i]; dx[i] = g.dat_dx[i]; w[i] = g.dat_w[i]; lo[i] = g.dat_lo[i];
  }
  s = g;
  s.node_t = nt; s.dat_ys = ys; s.dat_c = c; s.dat_dx = dx; s.dat_w = w; s.dat_lo = lo;
  __syncthreads();
  return true;
}

// Where a finished evaluation goes: the caller's arrays, or the accept step of the stretch move.
struct Sink {
  double* lnp;     // [W]
  int* status;     // [W] or null
  int* n_rhs;      // [W] or null
};

template <bool MOVE>
__device__ __forceinline__ void deliver(const Problem& p, const Work& k, const Sink& s, const Move& m, int sl, double lp_new,
                                        int st) {
  const int nr = k.n_rhs[sl];
  const int i = k.slot_wid ? k.slot_wid[sl] : sl;              // the walker this slot holds
  if (!MOVE) {
    s.lnp[i] = lp_new;
    if (s.status) s.status[i] = st;
    if (s.n_rhs) s.n_rhs[i] = nr;
    return;
  }
  const int ndim = p.ndim;
  const Draw d = draw_for(m, i);
  const int me = d.me;
  const double* q = m.prop + (size_t)i * ndim;
  const double lp_old = m.lnp[me];
  const double lnpdiff = __dadd_rn(__dadd_rn(__dmul_rn(ndim - 1.0, log(d.z)), lp_new), -lp_old);
  const bool accept = lnpdiff > log(d.ua);
  if (accept) {
    for (int c = 0; c < ndim; ++c) m.coords[(size_t)me * ndim + c] = q[c];
    m.lnp[me] = lp_new;
    if (m.accepted) m.accepted[me] += 1;
  }
  if (m.pack_out) {
    double* rowp = m.pack_out + (size_t)i * (ndim + 1);
    for (int c = 0; c < ndim; ++c) rowp[c] = accept ? q[c] : m.coords[(size_t)me * ndim + c];
    rowp[ndim] = accept ? lp_new : lp_old;
  }
  if (m.status) m.status[me] = st;
  if (m.bad_count && (st & (kWalkerIntegratorFail | kWalkerNonfiniteLnlike))) {
    const int slot = atomicAdd(m.bad_count, 1);
    if (slot < m.bad_capacity)
      for (int c = 0; c < ndim; ++c) m.bad_rows[(size_t)slot * ndim + c] = q[c];
  }
  if (m.n_rhs) m.n_rhs[me] = nr;
}

// One thread per walker, nodes one after the other: every lane of a warp walks the same node and datum
// indices (the dataset is shared), so the loop is converged.  Used when the dataset is small (synthetic
// light curves: 50 points) -- there a warp per walker would leave half its lanes without a node.
template <int MODE, bool MOVE>
__global__ void __launch_bounds__(64, 12) reduce_rows_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k,
                                                         const __grid_constant__ Sink s, double* __restrict__ out,
                                                         const __grid_constant__ Move m) {
  __shared__ double s_data[kDataSmemDoubles];
  DataView dv = p.dv;
  stage_data(p.dv, dv, s_data, 64);
  const int sl = blockIdx.x * 64 + threadIdx.x;
  if (sl >= k.W) return;
  int st = k.status[sl];
  double result = -INFINITY;
  if (!(st & kWalkerPriorReject)) {
    Walker w;
    rec_load_lum(k, sl, w);
    double* o = (MODE == kModeModelAtData) ? out + (size_t)(k.slot_wid ? k.slot_wid[sl] : sl) * dv.n_data : nullptr;
    const double chi2 = reduce_rows<MODE>(p.sp, dv, w, k.ybuf + (size_t)sl * k.ws, o, nullptr, 1, p.dat_orig, k.ns);
    if (MODE == kModeLnprob) result = lnlike_of(chi2, st);                       // + lnprior == 0.0
  }
  if (MODE == kModeLnprob) deliver<MOVE>(p, k, s, m, sl, result, st);
  else if (s.status) s.status[k.slot_wid ? k.slot_wid[sl] : sl] = st;
}

// One warp per walker, one DATUM per lane: the lane evaluates the luminosity stage at its datum's two
// bracketing nodes (one if the datum sits on a node), interpolates, forms its residual -- and the chi-square
// is a warp-shuffle reduction over the data.  Used for the large datasets (the short-GRB sample: up to 1944
// points, ~2 nodes per datum), and it is what makes a small ensemble on such a burst run at the speed of the
// integration alone instead of one thread walking thousands of nodes.
// ORDERED = true: the residuals are summed in datum order (passed round the warp by shuffles) instead of by a tree,
// which reproduces reduce_rows_kernel's chi-square bit for bit: small launches on SMALL datasets use this form --
// a 128-walker half-step then spends 5 us here instead of 80 -- while large ones keep one thread per walker, and a
// walker's lnprob still does not depend on the size of the batch it is in.
template <int MODE, bool MOVE, bool ORDERED>
__global__ void __launch_bounds__(128) reduce_coop_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k,
                                                          const __grid_constant__ Sink s, double* __restrict__ out,
                                                          const __grid_constant__ Move m) {
  const int lane = threadIdx.x & 31;
  const int sl = (blockIdx.x * 128 + threadIdx.x) >> 5;
  if (sl >= k.W) return;
  const int i = k.slot_wid ? k.slot_wid[sl] : sl;              // the walker this slot holds
  int st = k.status[sl];
  double chi2 = 0.0;
  if (!(st & kWalkerPriorReject)) {
    Walker w;
    rec_load_lum(k, sl, w);                                    // (every lane the same words: broadcast loads)
    const double* row = k.ybuf + (size_t)sl * k.ws;
    const DataView& dv = p.dv;
    for (int d0 = 0; d0 < dv.n_data; d0 += 32) {
      const int d = d0 + lane;
      double r = 0.0;
      if (d < dv.n_data) {
        const int lo = dv.dat_lo[d];
        const double dx = dv.dat_dx[d];
        double M, om;
        const double L_lo = node_luminosity(p.sp, w, dv.node_t[lo], dv.t_start, row[lo * k.ns], false, M, om).tot;
        double mod = L_lo;                                     // datum sits on a grid node
        if (dx != 0.0) {
          const double L_hi = node_luminosity(p.sp, w, dv.node_t[lo + 1], dv.t_start, row[(lo + 1) * k.ns], false, M, om).tot;
          mod = fma(L_hi - L_lo, dv.dat_w[d], L_lo);           // np.interp: slope*(x-x_lo)+y_lo
        }
        if (MODE == kModeLnprob) r = fma(-mod, dv.dat_c[d], dv.dat_ys[d]);   // (y - mod/1e50)/yerr
        else out[(size_t)i * dv.n_data + (p.dat_orig ? p.dat_orig[d] : d)] = mod * 1.0e-50;
      }
      if (MODE == kModeLnprob) {
        if (ORDERED) {
          const int cnt = min(32, dv.n_data - d0);
          for (int l = 0; l < cnt; ++l) {
            const double rl = __shfl_sync(kFull, r, l);
            chi2 = fma(rl, rl, chi2);
          }
        } else {
          chi2 = fma(r, r, chi2);
        }
      }
    }
    if (MODE == kModeLnprob && !ORDERED)
      for (int off = 16; off > 0; off >>= 1) chi2 += __shfl_xor_sync(kFull, chi2, off);   // same value on every lane
  }
  if (lane != 0) return;
  if (MODE == kModeLnprob) {
    double result = -INFINITY;
    if (!(st & kWalkerPriorReject)) result = lnlike_of(chi2, st);
    deliver<MOVE>(p, k, s, m, sl, result, st);
  } else if (s.status) {
    s.status[i] = st;
  }
}

// Light curves: one warp per walker, one NODE per lane, so the three output rows (and the state rows) are
// written as 256-byte runs along the node axis.  out [W][3][Nn] = Ltot, Lprop, Ldip (/1e50); state [W][2][Nn].
__global__ void __launch_bounds__(128) reduce_curves_kernel(const __grid_constant__ Problem p, const __grid_constant__ Work k,
                                                            double* __restrict__ out, double* __restrict__ state,
                                                            int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int sl = (blockIdx.x * 128 + threadIdx.x) >> 5;
  if (sl >= k.W) return;
  const int i = k.slot_wid ? k.slot_wid[sl] : sl;              // the walker this slot holds
  Walker w;
  rec_load_lum(k, sl, w);
  const int Nn = p.dv.n_nodes;
  const double* row = k.ybuf + (size_t)sl * k.ws;              // node-minor: k.ns == 1
  double* o = out + (size_t)i * 3 * Nn;
  double* so = state ? state + (size_t)i * 2 * Nn : nullptr;
  for (int j = lane; j < Nn; j += 32) {
    double M, om;
    const Lum L = node_luminosity(p.sp, w, p.dv.node_t[j], p.dv.t_start, row[j], so != nullptr, M, om);
    o[j] = L.tot * 1.0e-50;            // (/1e50, funcs.py:231,236, as one multiplication: <= 1 ulp)
    o[Nn + j] = L.prop * 1.0e-50;
    o[2 * Nn + j] = L.dip * 1.0e-50;
    if (so) {
      so[j] = M;
      so[Nn + j] = om;
    }
  }
  if (lane == 0 && status) status[i] = k.status[sl];
}

// Scatter of all-gathered packs (the collective exchange): packed row i of the half -> walker P(pos0 + i).
__global__ void unpack_kernel(SplitPerm perm, int pos0, int n, int ndim, const double* __restrict__ packed,
                              double* __restrict__ coords, double* __restrict__ lnp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t w = perm_at(perm, (uint32_t)(pos0 + i));
  const double* row = packed + (size_t)i * (ndim + 1);
  for (int d = 0; d < ndim; ++d) coords[w * ndim + d] = row[d];
  lnp[w] = row[ndim];
}

// Bring this rank's replica up to date (peer-mapped ensembles): every row whose last mover was another rank is
// fetched from that rank's replica.  `m.prev_perm` is the order of the last completed step.
__global__ void sync_replica_kernel(const __grid_constant__ Move m, int n, int ndim) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const int home = home_before_this_step(m, w);
  if (home == m.rank) return;
  const double* src = replica_coords(m, home);
  for (int c = 0; c < ndim; ++c) m.coords[(size_t)w * ndim + c] = src[(size_t)w * ndim + c];
  m.lnp[w] = replica_lnp(m, home)[w];
}

__global__ void order_kernel(SplitPerm perm, int n, int* __restrict__ order) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) order[g] = (int)perm_at(perm, (uint32_t)g);
}

// ---- cross-GPU flag barrier over peer-mapped memory ---------------------------------------------
// Thread p of the one block: raise flag[rank] = epoch in peer p's array (a system-scope release, after the
// preceding kernel's peer stores -- stream order plus the fence make them visible first), then spin on
// this rank's own array until peer p has raised its flag (system-scope acquire).  The spin is on LOCAL
// memory; the only NVLink traffic is one 8-byte store per peer.
struct PeerFlags {
  uint64_t* p[MP_MAX_PEERS + 1];
};
__global__ void peer_barrier_kernel(uint64_t* my_flags, PeerFlags pf, int rank, int world, uint64_t epoch, int* error) {
  const int p = threadIdx.x;
  if (p >= world || p == rank) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pf.p[p] + rank), "l"(epoch) : "memory");
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + p) : "memory");
    if (v >= epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) {   // 10 s: a peer died -- report instead of hanging the GPU
      if (error) *error = 1;
      break;
    }
    __nanosleep(200);
  }
}

// ---- posterior summaries of a device-resident chain (plot_synth.py:150-166) -------------------------------
// Column means and the centred cross-products (for np.corrcoef), and exact order statistics by radix select
// (for np.percentile): the chain never leaves the device, the host receives a few dozen numbers.
__global__ void chain_mean_kernel(const double* __restrict__ x, long long n, int ndim, double* __restrict__ sums) {
  __shared__ double sh[MP_MAX_NDIM][8];
  double acc[MP_MAX_NDIM];
  for (int d = 0; d < MP_MAX_NDIM; ++d) acc[d] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int d = 0; d < ndim; ++d) acc[d] += x[i * ndim + d];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int d = 0; d < ndim; ++d) {
    double v = acc[d];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    if (lane == 0) sh[d][wrp] = v;
  }
  __syncthreads();
  if (threadIdx.x < ndim) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sh[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, v);
  }
}

__global__ void chain_cov_kernel(const double* __restrict__ x, long long n, int ndim, const double* __restrict__ mean,
                                 double* __restrict__ cov /*[ndim][ndim], upper triangle*/) {
  __shared__ double sh[MP_MAX_NDIM * MP_MAX_NDIM][8];
  double acc[MP_MAX_NDIM * (MP_MAX_NDIM + 1) / 2];
  const int npair = ndim * (ndim + 1) / 2;
  for (int q = 0; q < npair; ++q) acc[q] = 0.0;
  double mu[MP_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) mu[d] = mean[d];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double c[MP_MAX_NDIM];
    for (int d = 0; d < ndim; ++d) c[d] = x[i * ndim + d] - mu[d];
    int q = 0;
    for (int a = 0; a < ndim; ++a)
      for (int b = a; b < ndim; ++b, ++q) acc[q] = fma(c[a], c[b], acc[q]);
  }
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int q = 0; q < npair; ++q) {
    double v = acc[q];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    if (lane == 0) sh[q][wrp] = v;
  }
  __syncthreads();
  if (threadIdx.x < npair) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sh[threadIdx.x][w];
    int q = 0;
    for (int a = 0; a < ndim; ++a)
      for (int b = a; b < ndim; ++b, ++q)
        if (q == (int)threadIdx.x) atomicAdd(cov + a * ndim + b, v);
  }
}

// IEEE doubles as unsigned keys in ascending order (NaNs last).
__device__ __forceinline__ unsigned long long order_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// One pass of the radix select on column `col`: histogram of the 16-bit digit at `shift` over the elements whose
// higher digits equal those of `prefix`.
__global__ void select_hist_kernel(const double* __restrict__ x, long long n, int ndim, int col, unsigned long long prefix,
                                   int shift, unsigned* __restrict__ hist) {
  const unsigned long long himask = (shift >= 48) ? 0ull : (~0ull << (shift + 16));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = order_key(x[i * ndim + col]);
    if ((key & himask) == (prefix & himask)) atomicAdd(hist + (unsigned)((key >> shift) & 0xffffull), 1u);
  }
}

// ---- the coupled right-hand side, as ODEs()/odes() return it -----------------------
// funcs.py:75-142 / magnetar/funcs.py:33-101, operation order kept.
__global__ void rhs_kernel(double inertia_factor, double mdot_factor, double breakup, int dipole_torque,
                           const double* __restrict__ y, const double* __restrict__ t,
                           const double* __restrict__ pars, double n, double alpha, double cs7, double k,
                           int W, double* __restrict__ dydt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const double Mdisc = y[2 * i], omega = y[2 * i + 1], tt = t[i];
  const double B = pars[5 * i], MdiscI = pars[5 * i + 1], RdiscI = pars[5 * i + 2];
  const double epsilon = pars[5 * i + 3], delta = pars[5 * i + 4];
  const double inertia = inertia_factor * kM * (kR * kR);
  const double Rdisc = RdiscI * 1.0e5;
  const double tvisc = Rdisc / (alpha * cs7 * 1.0e7);
  const double mu = 1.0e15 * B * (kR * kR * kR);
  const double M0 = delta * MdiscI * kMsol;
  const double tfb = epsilon * tvisc;
  double Rm = pow(mu, 4.0 / 7.0) * pow(kGM, -1.0 / 7.0) * pow((mdot_factor * Mdisc) / tvisc, -2.0 / 7.0);
  const double Rc = pow(kGM / (omega * omega), 1.0 / 3.0);
  const double Rlc = kC / omega;
  if (Rm >= k * Rlc) Rm = k * Rlc;
  const double w = pow(Rm / Rc, 1.5);
  const double x = kGM / (kR * (kC * kC));
  const double modW = 0.6 * kM * (kC * kC) * (x / (1.0 - 0.5 * x));
  const double rot_param = (0.5 * inertia * (omega * omega)) / modW;
  double Ndip = (-1.0 * (mu * mu) * (omega * omega * omega)) / (6.0 * (kC * kC * kC));
  if (dipole_torque) {                                  // figure_3.py:142-143
    const double q = Rlc / Rm;
    Ndip = (-2.0 / 3.0) * (((mu * mu) * (omega * omega * omega)) / (kC * kC * kC)) * (q * q * q);
  }
  const double eta2 = 0.5 * (1.0 + tanh(n * (w - 1.0)));
  const double eta1 = 1.0 - eta2;
  const double Mdotprop = eta2 * (Mdisc / tvisc);
  const double Mdotacc = eta1 * (Mdisc / tvisc);
  const double Mdotfb = (M0 / tfb) * pow((tt + tfb) / tfb, -5.0 / 3.0);
  double Nacc;
  if (rot_param > breakup) Nacc = 0.0;
  else if (Rm >= kR) Nacc = sqrt(kGM * Rm) * (Mdotacc - Mdotprop);
  else Nacc = sqrt(kGM * kR) * (Mdotacc - Mdotprop);
  dydt[2 * i] = Mdotfb - Mdotacc - Mdotprop;
  dydt[2 * i + 1] = (Nacc + Ndip) / inertia;
}

// ---- the comparison model of code/figure_5.py:222-363 ("Ben's model") --------------------------------------
// An accreting-mass magnetar with an exponentially draining disc, advanced by explicit Euler steps of 1 s (the
// reference: a Python loop over 1e6 array elements per parameter set).  One parameter set per thread, the script's
// operation order kept (fractional powers through pow, as NumPy's **); every `stride`-th step is stored.
// out [W][3][n_out] = Ltot, Lprop, Ldip in units of 1e50 erg/s (figure_5.py:354-368).
__global__ void gompertz_kernel(const double* __restrict__ pars, int W, double alpha, double cs7, double k, double omass,
                                double dipeff, double propeff, long long n_steps, int stride, long long n_out,
                                double* __restrict__ out) {
  const int wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= W) return;
  const double B = pars[6 * wi], P = pars[6 * wi + 1], MdiscI = pars[6 * wi + 2], RdiscI = pars[6 * wi + 3];
  const double spin = P * 1.0e-3;                               // :224
  const double Rdisc = RdiscI * 1.0e5;
  const double visc = alpha * cs7 * 1.0e7 * Rdisc;
  const double mu = 1.0e15 * B * (kR * kR * kR);
  double omega = (2.0 * 3.141592653589793) / spin;             // :229
  const double Mdisc0 = MdiscI * kMsol;
  double M_bg = omass * kMsol;
  const double Mdot0 = (3.0 * Mdisc0 * visc) / (Rdisc * Rdisc);  // :258
  double Mdot = Mdot0, Msum = 0.0, tt = 1.0, omegadot = 0.0;
  const double mu47 = pow(mu, 4.0 / 7.0), mu2 = mu * mu, c3 = kC * kC * kC, c2 = kC * kC, R2 = kR * kR, R3 = kR * kR * kR;
  double* o = out + (size_t)wi * 3 * n_out;
  for (long long i = 0; i < n_steps; ++i) {
    if (i > 0) {                                                // :302-309
      tt = tt + 1.0;
      omega = omega + omegadot;
      M_bg = M_bg + Msum;
      Mdot = Mdot0 * exp((-3.0 * visc * tt) / (Rdisc * Rdisc));
    }
    const double GMb = kG * M_bg;
    double Rm = mu47 * pow(GMb, -1.0 / 7.0) * pow(Mdot, -2.0 / 7.0);
    const double Rc = pow(GMb / (omega * omega), 1.0 / 3.0);
    const double light = kC / omega;
    if (Rm >= (k * light)) Rm = k * light;
    const double lr = light / Rm;
    const double Ndip = (-2.0 / 3.0) * ((mu2 * (omega * omega * omega)) / c3) * (lr * lr * lr);
    const double w = pow(Rm / Rc, 1.5);
    const double nn = 1.0 - w;
    const double inertia = 0.35 * M_bg * R2;
    const double bigT = 0.5 * inertia * (omega * omega);
    const double x = GMb / (kR * c2);
    const double modW = 0.6 * M_bg * c2 * (x / (1.0 - 0.5 * x));
    const double beta = bigT / modW;
    double Nacc;
    if (beta > 0.27) {
      Nacc = 0.0;
    } else if (Rm >= kR) {
      Nacc = nn * sqrt(GMb * Rm) * Mdot;
      if (!isfinite(Nacc)) Nacc = 0.0;
    } else {
      const double f = 1.0 - (omega / sqrt(GMb / R3));
      Nacc = (i == 0) ? f / sqrt(GMb * kR) * Mdot : f * sqrt(GMb * kR) * Mdot;   // (:283-285 divides, :335-337 multiplies)
      if (!isfinite(Nacc)) Nacc = 0.0;
    }
    if (Rc >= Rm) Msum = Mdot;                                  // :293-296, :341-344
    else if (i == 0) Msum = 0.0;
    omegadot = (Ndip + Nacc) / inertia;
    if (i % stride == 0) {
      double lp = (-1.0 * Nacc * omega) - ((GMb * Mdot) / Rm);
      double ld = (mu2 * ((omega * omega) * (omega * omega))) / (6.0 * c3);
      if (!isfinite(lp) || lp <= 0.0) lp = 0.0;                 // :354-361
      if (!isfinite(ld) || ld <= 0.0) ld = 0.0;
      const long long jo = i / stride;
      o[jo] = ((propeff * lp) + (dipeff * ld)) * 1.0e-50;
      o[n_out + jo] = lp * 1.0e-50;
      o[2 * n_out + jo] = ld * 1.0e-50;
    }
  }
}

// ---- FP64 FMA peak: 8 independent chains per thread, 2 flop per DFMA ---------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.9999999, c = 1e-9;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

}  // namespace mp

// =====================================================================================
//                                     C ABI
// =====================================================================================
using namespace mp;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define MP_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(MP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

struct DeviceNodes {
  double* node_t = nullptr;
  int n_nodes = 0;
};

struct mp_handle {
  int device = 0;
  mp_model_spec model;
  Spec spec;
  mp_prior_spec prior;
  std::vector<double> grid;
  // data
  int D = 0;
  NodeProgram np;
  DeviceNodes data_nodes;
  double *d_y = nullptr, *d_yerr = nullptr, *d_dx = nullptr, *d_Dx = nullptr;   // y/yerr, 1e-50/yerr, dx, dx/Dx (DataView)
  int *d_lo = nullptr, *d_orig = nullptr;
  // curve node sets per stride
  std::map<int, DeviceNodes> curve_nodes;
  // staging for the host-pointer entry points
  double *s_theta = nullptr, *s_out = nullptr, *s_state = nullptr, *s_lnp = nullptr;
  int *s_status = nullptr, *s_nrhs = nullptr, *s_cstatus = nullptr;
  size_t cap_theta = 0, cap_out = 0, cap_state = 0, cap_w = 0, cap_cstatus = 0;
  // A lane = a stream plus the work space the stages of a launch hand to one another.  The handle owns two
  // pipeline lanes so that the host-pointer entry points can overlap the H2D copy of one chunk with the
  // kernels of the previous one.  Device-pointer entry points run on the caller's stream, and every such
  // stream gets a lane of its own (work space only): calls on one handle from different streams -- several
  // ensembles on one dataset, a device call next to an mp_lnprob_batch_async in flight -- never share work
  // space; calls on ONE stream are ordered by the stream.
  struct Lane {
    cudaStream_t stream = nullptr;
    unsigned long long* recs = nullptr;
    double* ybuf = nullptr;
    int *status = nullptr, *n_rhs = nullptr, *work = nullptr, *counters = nullptr, *key = nullptr, *hist = nullptr,
        *slot_wid = nullptr;
    StiffRec* squeue = nullptr;
    double* prop = nullptr;
    size_t cap_walkers = 0, cap_ybuf = 0, cap_prop = 0;
    int* hint_host = nullptr;    // mapped pinned int[4]: counters[4..7] of the last large launch on this lane (see launch_eval)
    int* hint_dev = nullptr;     // the device's pointer to it
  } lanes[2];
  std::map<cudaStream_t, Lane> user_lanes;
  std::mutex user_lanes_mu;
  int sm_count = 148;
  cudaStream_t stream = nullptr;   // == lanes[0].stream
  Lane* last_lane = nullptr;       // work space of the most recent launch (mp_last_stiff_count)
  int64_t kernels_launched = 0;    // evaluation-pipeline kernels launched through this handle so far (mp_kernels_launched)
};

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return MP_OK;
  MP_CUDA(cudaMalloc((void**)dst, v.size() * sizeof(T)));
  MP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MP_OK;
}

// Grow-only device buffer.  (cudaFree waits for the device, so a buffer an earlier launch still uses is
// released only after that launch has finished.)
template <typename T>
static int ensure(T** p, size_t* cap, size_t need) {
  if (need <= *cap) return MP_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  MP_CUDA(cudaMalloc((void**)p, need * sizeof(T)));
  *cap = need;
  return MP_OK;
}

extern "C" int mp_abi_version(void) { return MP_ABI_VERSION; }

extern "C" int mp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char* mp_last_error(void) { return g_err.c_str(); }

extern "C" int mp_create(const mp_model_spec* spec, const mp_prior_spec* prior, const double* grid,
                         int32_t G, const double* t, const double* y, const double* yerr, int32_t D,
                         int32_t device, mp_handle** out) {
  if (!spec || !grid || !out || D < 0 || (D > 0 && (!t || !y || !yerr)))
    return fail(MP_ERR_BAD_ARG, "mp_create: null pointer or negative size");
  *out = nullptr;
  NodeProgram np;
  int rc = build_node_program(grid, G, t, y, yerr, D, np);
  if (rc == MP_ERR_BAD_GRID) return fail(rc, "mp_create: grid must be strictly increasing with >= 2 nodes");
  if (rc == MP_ERR_DATA_RANGE)
    return fail(rc, "A value in x_new is outside the interpolation range.");
  if (mp_device_count() <= device || device < 0)
    return fail(MP_ERR_CUDA, "mp_create: no such CUDA device (magprop_b200 has no CPU path)");
  MP_CUDA(cudaSetDevice(device));
  mp_handle* h = new mp_handle();
  h->device = device;
  h->model = *spec;
  h->spec = make_spec(*spec);
  if (prior) h->prior = *prior;
  else std::memset(&h->prior, 0, sizeof(h->prior));
  h->grid.assign(grid, grid + G);
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  }
  for (auto& L : h->lanes) {
    if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) {
      mp_destroy(h);
      return fail(MP_ERR_CUDA, "mp_create: stream creation failed");
    }
  }
  h->stream = h->lanes[0].stream;
  h->D = D;
  h->np = np;
  h->data_nodes.n_nodes = (int)np.node_t.size();
  if ((rc = upload(&h->data_nodes.node_t, np.node_t)) || (rc = upload(&h->d_y, np.ys)) ||
      (rc = upload(&h->d_yerr, np.c)) || (rc = upload(&h->d_dx, np.dx)) ||
      (rc = upload(&h->d_Dx, np.w)) || (rc = upload(&h->d_lo, np.lo)) ||
      (rc = upload(&h->d_orig, np.order))) {
    mp_destroy(h);
    return rc;
  }
  *out = h;
  return MP_OK;
}

extern "C" void mp_destroy(mp_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->data_nodes.node_t);
  cudaFree(h->d_y); cudaFree(h->d_yerr); cudaFree(h->d_dx); cudaFree(h->d_Dx);
  cudaFree(h->d_lo); cudaFree(h->d_orig);
  for (auto& kv : h->curve_nodes) cudaFree(kv.second.node_t);
  cudaFree(h->s_theta); cudaFree(h->s_out); cudaFree(h->s_state); cudaFree(h->s_lnp);
  cudaFree(h->s_status); cudaFree(h->s_nrhs); cudaFree(h->s_cstatus);
  auto free_lane = [](mp_handle::Lane& L) {
    if (L.hint_host) cudaFreeHost(L.hint_host);
    cudaFree(L.recs); cudaFree(L.ybuf); cudaFree(L.status); cudaFree(L.n_rhs); cudaFree(L.work);
    cudaFree(L.counters); cudaFree(L.squeue); cudaFree(L.prop); cudaFree(L.key); cudaFree(L.hist); cudaFree(L.slot_wid);
    if (L.stream) cudaStreamDestroy(L.stream);
  };
  for (auto& L : h->lanes) free_lane(L);
  for (auto& kv : h->user_lanes) free_lane(kv.second);     // (their .stream is null: the streams are the callers')
  delete h;
}

extern "C" int mp_set_prior(mp_handle* h, const mp_prior_spec* prior) {
  if (!h || !prior) return fail(MP_ERR_BAD_ARG, "mp_set_prior: null pointer");
  h->prior = *prior;
  return MP_OK;
}

static int fill_problem(mp_handle* h, Problem& p, const DeviceNodes& nodes, bool with_data, int ndim, int W,
                        bool use_prior) {
  if (ndim < 6 || ndim > MP_MAX_NDIM) return fail(MP_ERR_BAD_ARG, "ndim must be 6, 7, 8 or 9");
  if (W < 0) return fail(MP_ERR_BAD_ARG, "negative walker count");
  if (use_prior && h->prior.enabled && h->prior.ndim != ndim)
    return fail(MP_ERR_BAD_ARG, "prior dimension does not match ndim");
  std::memset(&p, 0, sizeof(p));
  p.sp = h->spec;
  p.dv.n_nodes = nodes.n_nodes;
  p.dv.node_t = nodes.node_t;
  p.dv.t_start = h->grid[0];
  if (with_data) {
    p.dv.n_data = h->D;
    p.dv.dat_ys = h->d_y;
    p.dv.dat_c = h->d_yerr;
    p.dv.dat_dx = h->d_dx;
    p.dv.dat_w = h->d_Dx;
    p.dv.dat_lo = h->d_lo;
    p.dat_orig = h->d_orig;
  }
  p.prior_enabled = use_prior ? h->prior.enabled : 0;
  for (int i = 0; i < MP_MAX_NDIM; ++i) {
    p.lower[i] = h->prior.lower[i];
    p.upper[i] = h->prior.upper[i];
  }
  p.ndim = ndim;
  return MP_OK;
}

// lane >= 0: one of the handle's own pipeline lanes; lane < 0: the lane of the caller's stream
static int lane_for(mp_handle* h, cudaStream_t stream, int lane, mp_handle::Lane** out) {
  if (lane >= 0) {
    *out = &h->lanes[lane];
  } else {
    std::lock_guard<std::mutex> g(h->user_lanes_mu);
    *out = &h->user_lanes[stream];
  }
  h->last_lane = *out;
  return MP_OK;
}

// Walkers per slab: a launch of more walkers than this runs as several slabs one after the other, so the work
// space stays below ~1.5 GiB whatever the ensemble (10^7 walkers) or the node count (10 001 for full curves).
static int slab_walkers(int Nn) {
  const size_t per_walker = sizeof(WalkerRec) + sizeof(StiffRec) + 8 * (size_t)Nn + 16;
  const size_t n = ((size_t)3 << 29) / per_walker;
  return (int)std::min<size_t>(std::max<size_t>(n, 4096), (size_t)1 << 22);
}

static int ensure_work(mp_handle::Lane& L, int S, int Nn, int ndim_prop, Work& k) {
  int rc = MP_OK;
  if ((size_t)S > L.cap_walkers) {
    cudaFree(L.recs); cudaFree(L.status); cudaFree(L.n_rhs); cudaFree(L.work); cudaFree(L.squeue); cudaFree(L.key); cudaFree(L.slot_wid);
    L.recs = nullptr; L.status = L.n_rhs = L.work = L.key = L.slot_wid = nullptr; L.squeue = nullptr; L.cap_walkers = 0;
    MP_CUDA(cudaMalloc((void**)&L.recs, (size_t)S * sizeof(WalkerRec)));      // kRecWords rows of S words
    MP_CUDA(cudaMalloc((void**)&L.status, (size_t)S * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.n_rhs, (size_t)S * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.work, (size_t)S * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.key, (size_t)S * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.slot_wid, (size_t)S * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.squeue, (size_t)S * sizeof(StiffRec)));
    L.cap_walkers = (size_t)S;
  }
  if (!L.counters) {             // [8..11]: what the host was told last (never reset)
    MP_CUDA(cudaMalloc((void**)&L.counters, 12 * sizeof(int)));
    MP_CUDA(cudaMemset(L.counters, 0, 12 * sizeof(int)));
  }
  if (!L.hist) MP_CUDA(cudaMalloc((void**)&L.hist, (kOrderBuckets + 3) * sizeof(int)));
  if (!L.hint_host && cudaHostAlloc((void**)&L.hint_host, 4 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    std::memset(L.hint_host, 0, 4 * sizeof(int));
    if (cudaHostGetDevicePointer((void**)&L.hint_dev, L.hint_host, 0) != cudaSuccess) L.hint_dev = nullptr;
  }
  if ((rc = ensure(&L.ybuf, &L.cap_ybuf, (size_t)S * Nn))) return rc;
  if (ndim_prop > 0 && (rc = ensure(&L.prop, &L.cap_prop, (size_t)S * ndim_prop))) return rc;
  k.stride = (int)L.cap_walkers;
  k.recs = L.recs; k.ybuf = L.ybuf; k.status = L.status; k.n_rhs = L.n_rhs; k.work = L.work;
  k.squeue = L.squeue; k.counters = L.counters;
  return MP_OK;
}

// One evaluation of W walkers: setup -> advance (explicit, then the stiff queue) -> reduce, slab by slab.
//   MOVE = false: parameters from d_theta [W][ndim]; results to `sink` (lnprob), `out` (model at data
//                 [W][D], or curves [W][3][Nn]) and `state` (curves, optional)
//   MOVE = true : the W movers of a stretch-move half-step described by `mv`
template <int MODE, bool MOVE>
static int launch_eval(mp_handle* h, const Problem& p, const double* d_theta, int W, Sink sink, double* out, double* state,
                       const Move* mv, cudaStream_t stream, int lane = -1) {
  if (W == 0) return MP_OK;
  mp_handle::Lane* Lp = nullptr;
  lane_for(h, stream, lane, &Lp);
  const int Nn = p.dv.n_nodes, D = p.dv.n_data, ndim = p.ndim;
  const int S = std::min(W, slab_walkers(Nn));
  Work k;
  std::memset(&k, 0, sizeof(k));
  int rc = ensure_work(*Lp, S, Nn, MOVE ? ndim : 0, k);
  if (rc) return rc;
  Move m;
  if (MOVE) m = *mv;
  else std::memset(&m, 0, sizeof(m));
  const bool coop = Nn > 64 || MODE == kModeCurves;   // by the DATASET only: a walker's lnprob never depends on the batch it is in
  k.ws = coop ? (size_t)Nn : 1;
  k.ns = coop ? 1 : (size_t)k.stride;
  for (int i0 = 0; i0 < W; i0 += S) {
    const int n = std::min(S, W - i0);
    k.W = n;
    Move ms = m;
    if (MOVE) {
      if (ms.active) ms.active += i0;
      else ms.pos0 += i0;
      ms.prop = Lp->prop;
      if (ms.pack_out) ms.pack_out += (size_t)i0 * (ndim + 1);
    }
    Sink sk = sink;
    if (sk.lnp) sk.lnp += i0;
    if (sk.status) sk.status += i0;
    if (sk.n_rhs) sk.n_rhs += i0;
    MP_CUDA(cudaMemsetAsync(k.counters, 0, 8 * sizeof(int), stream));
    // (a sharded move that reads its rows from peer replicas forms its proposals inside the setup kernel, where the
    // remote reads hide behind the other threads' arithmetic: in a kernel of their own they cost 0.07 ms per half-step)
    // A tight ensemble gains nothing from the ordering and pays two launches for it.  How far an ensemble reaches in the
    // ordering's bins is measured by every large launch's setup kernel; the four numbers travel to pinned host memory
    // behind the launch and are read HERE, unsynchronised, by the next launch on the same lane (a chain's ensemble moves
    // slowly): inside kTightMdBins x kTightEpsBins the walkers are taken as they come.  Results do not depend on it (a
    // walker's arithmetic is the same in any slot).
    bool ordered = n >= kOrderMinWalkers && Nn > 0 && !(MOVE && m.n_peers > 0);
    k.track = (ordered && Lp->hint_dev) ? 1 : 0;
    k.hint = Lp->hint_dev;
    if (ordered && Lp->hint_host) {
      const volatile int* e = Lp->hint_host;
      const int e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
      if (e0 > 0 && e0 + e1 - 257 <= kTightMdBins && e2 + e3 - 33 <= kTightEpsBins) ordered = false;
    }
    const double* th0 = MOVE ? nullptr : d_theta + (size_t)i0 * ndim;
    k.key = Lp->key;
    k.hist = Lp->hist;
    k.slot_wid = ordered ? Lp->slot_wid : nullptr;
    if (ordered) {
      // slots in key order: keys + histogram (MOVE: and the proposals), scan, scatter -- then the setup runs per slot
      MP_CUDA(cudaMemsetAsync(k.hist, 0, (kOrderBuckets + 3) * sizeof(int), stream));
      order_key_kernel<MOVE><<<(n + 127) / 128, 128, 0, stream>>>(p, k, th0, ms);
      order_scatter_kernel<<<(n + 1023) / 1024, 1024, 0, stream>>>(n, k.key, k.hist, k.slot_wid);
      h->kernels_launched += 2;
    }
    setup_kernel<MOVE><<<(n + 127) / 128, 128, 0, stream>>>(p, k, th0, ms);
    // small launches: 32-thread blocks spread the warps over more SMs
    if (Nn == 0) {
      // (a handle without data -- lnprior-only callers: nothing to integrate, lnlike = -0.5 * 0)
    } else if (n <= h->sm_count * 64 * 4) {
      advance_kernel<false, 32><<<std::min((n + 31) / 32, h->sm_count * MP_MIN_BLOCKS_32), 32, 0, stream>>>(p, k);
      advance_kernel<true, 32><<<std::min((n + 31) / 32, h->sm_count * MP_STIFF_MIN_WARPS), 32, 0, stream>>>(p, k);
    } else {
      advance_kernel<false, 64><<<std::min((n + 63) / 64, h->sm_count * MP_MIN_BLOCKS_64), 64, 0, stream>>>(p, k);
      if (n >= kStiffWideMinWalkers)
        advance_kernel<true, 64, kStiffWarpsWide><<<std::min((n + 63) / 64, h->sm_count * kStiffWarpsWide / 2), 64, 0, stream>>>(p, k);
      else
        advance_kernel<true, 64><<<std::min((n + 63) / 64, h->sm_count * MP_STIFF_MIN_WARPS / 2), 64, 0, stream>>>(p, k);
    }
    if (MODE == kModeCurves) {
      reduce_curves_kernel<<<(n + 3) / 4, 128, 0, stream>>>(p, k, out + (size_t)i0 * 3 * Nn,
                                                            state ? state + (size_t)i0 * 2 * Nn : nullptr, sk.status);
    } else {
      double* o = out ? out + (size_t)i0 * D : nullptr;
      if (coop) reduce_coop_kernel<MODE, MOVE, false><<<(n + 3) / 4, 128, 0, stream>>>(p, k, sk, o, ms);
      else if (n <= kCoopSmallWalkers) reduce_coop_kernel<MODE, MOVE, true><<<(n + 3) / 4, 128, 0, stream>>>(p, k, sk, o, ms);
      else reduce_rows_kernel<MODE, MOVE><<<(n + 63) / 64, 64, 0, stream>>>(p, k, sk, o, ms);
    }
    h->kernels_launched += (Nn == 0) ? 2 : 4;      // setup, advance (explicit), advance (implicit), reduce
    MP_CUDA(cudaGetLastError());
  }
  return MP_OK;
}

extern "C" int64_t mp_kernels_launched(const mp_handle* h) { return h ? h->kernels_launched : 0; }

extern "C" int mp_lnprob_batch_device(mp_handle* h, const double* d_theta, int32_t W, int32_t ndim,
                                      double* d_lnp, int32_t* d_status, int32_t* d_n_rhs, void* stream) {
  if (!h || (W > 0 && (!d_theta || !d_lnp))) return fail(MP_ERR_BAD_ARG, "mp_lnprob_batch_device: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  Problem p;
  int rc = fill_problem(h, p, h->data_nodes, true, ndim, W, true);
  if (rc) return rc;
  return launch_eval<kModeLnprob, false>(h, p, d_theta, W, Sink{d_lnp, d_status, d_n_rhs}, nullptr, nullptr, nullptr,
                                         (cudaStream_t)stream);
}

// One full wave of the explicit integrator: its resident 64-thread blocks on every SM.
static int wave_walkers(const mp_handle* h) { return h->sm_count * MP_MIN_BLOCKS_64 * 64; }

extern "C" int mp_lnprob_batch_async(mp_handle* h, const double* theta, int32_t W, int32_t ndim, double* lnp,
                                     int32_t* status, int32_t* n_rhs) {
  if (!h || (W > 0 && (!theta || !lnp))) return fail(MP_ERR_BAD_ARG, "mp_lnprob_batch: null pointer");
  if (ndim < 6 || ndim > MP_MAX_NDIM) return fail(MP_ERR_BAD_ARG, "ndim must be 6, 7, 8 or 9");
  if (W == 0) return MP_OK;
  MP_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure(&h->s_theta, &h->cap_theta, (size_t)W * MP_MAX_NDIM))) return rc;
  if (W > (int)h->cap_w) {
    cudaFree(h->s_lnp); cudaFree(h->s_status); cudaFree(h->s_nrhs);
    h->s_lnp = nullptr; h->s_status = nullptr; h->s_nrhs = nullptr; h->cap_w = 0;
    MP_CUDA(cudaMalloc((void**)&h->s_lnp, (size_t)W * sizeof(double)));
    MP_CUDA(cudaMalloc((void**)&h->s_status, (size_t)W * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&h->s_nrhs, (size_t)W * sizeof(int)));
    h->cap_w = W;
  }
  // Large batches go through in chunks of a few waves on two alternating lanes: the H2D copy of chunk
  // k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k (when the caller's buffers are
  // pinned; pageable buffers still work, the copies just serialise).
  const int wave = 4 * wave_walkers(h);
  const int chunk = (W <= wave + wave / 2) ? W : wave;
  Problem p;
  if ((rc = fill_problem(h, p, h->data_nodes, true, ndim, W, true))) return rc;
  int k = 0;
  for (int c0 = 0; c0 < W; c0 += chunk, ++k) {
    const int lane = k & 1;
    cudaStream_t st = h->lanes[lane].stream;
    int n = W - c0;
    if (n > chunk + chunk / 2) n = chunk;          // the last chunk absorbs a remainder below half a chunk
    MP_CUDA(cudaMemcpyAsync(h->s_theta + (size_t)c0 * ndim, theta + (size_t)c0 * ndim, (size_t)n * ndim * sizeof(double),
                            cudaMemcpyHostToDevice, st));
    if ((rc = launch_eval<kModeLnprob, false>(h, p, h->s_theta + (size_t)c0 * ndim, n,
                                              Sink{h->s_lnp + c0, h->s_status + c0, h->s_nrhs + c0}, nullptr, nullptr, nullptr,
                                              st, lane)))
      return rc;
    MP_CUDA(cudaMemcpyAsync(lnp + c0, h->s_lnp + c0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (status) MP_CUDA(cudaMemcpyAsync(status + c0, h->s_status + c0, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_rhs) MP_CUDA(cudaMemcpyAsync(n_rhs + c0, h->s_nrhs + c0, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n != chunk) break;
  }
  return MP_OK;
}

extern "C" int mp_synchronize(mp_handle* h) {
  if (!h) return fail(MP_ERR_BAD_ARG, "mp_synchronize: null handle");
  MP_CUDA(cudaSetDevice(h->device));
  MP_CUDA(cudaStreamSynchronize(h->lanes[0].stream));
  MP_CUDA(cudaStreamSynchronize(h->lanes[1].stream));
  return MP_OK;
}

extern "C" int mp_lnprob_batch(mp_handle* h, const double* theta, int32_t W, int32_t ndim, double* lnp,
                               int32_t* status, int32_t* n_rhs) {
  const int rc = mp_lnprob_batch_async(h, theta, W, ndim, lnp, status, n_rhs);
  if (rc) return rc;
  return (W > 0) ? mp_synchronize(h) : MP_OK;
}

extern "C" int mp_model_at_data(mp_handle* h, const double* pars, int32_t W, int32_t ndim, double* out,
                                int32_t* status) {
  if (!h || (W > 0 && (!pars || !out))) return fail(MP_ERR_BAD_ARG, "mp_model_at_data: null pointer");
  if (h->D == 0) return fail(MP_ERR_NO_DATA, "mp_model_at_data: handle has no data times");
  if (W == 0) return MP_OK;
  MP_CUDA(cudaSetDevice(h->device));
  Problem p;
  int rc = fill_problem(h, p, h->data_nodes, true, ndim, W, false);
  if (rc) return rc;
  p.sp.unlog_mask = 0;  // model_lum takes physical parameters
  if ((rc = ensure(&h->s_theta, &h->cap_theta, (size_t)W * MP_MAX_NDIM))) return rc;
  if ((rc = ensure(&h->s_out, &h->cap_out, (size_t)W * h->D))) return rc;
  if ((rc = ensure(&h->s_cstatus, &h->cap_cstatus, (size_t)W))) return rc;
  MP_CUDA(cudaMemcpyAsync(h->s_theta, pars, (size_t)W * ndim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = launch_eval<kModeModelAtData, false>(h, p, h->s_theta, W, Sink{nullptr, h->s_cstatus, nullptr}, h->s_out, nullptr,
                                                 nullptr, h->stream, 0)))
    return rc;
  MP_CUDA(cudaMemcpyAsync(out, h->s_out, (size_t)W * h->D * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (status) MP_CUDA(cudaMemcpyAsync(status, h->s_cstatus, (size_t)W * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MP_CUDA(cudaStreamSynchronize(h->stream));
  return MP_OK;
}

static int curve_node_set(mp_handle* h, int stride, DeviceNodes** out) {
  if (stride < 1) stride = 1;
  auto it = h->curve_nodes.find(stride);
  if (it == h->curve_nodes.end()) {
    std::vector<double> nt;
    std::vector<int> gi;
    build_curve_nodes(h->grid.data(), (int)h->grid.size(), stride, nt, gi);
    DeviceNodes dn;
    dn.n_nodes = (int)nt.size();
    int rc = upload(&dn.node_t, nt);
    if (rc) return rc;
    it = h->curve_nodes.emplace(stride, dn).first;
  }
  *out = &it->second;
  return MP_OK;
}

extern "C" int32_t mp_curve_nodes(const mp_handle* h, int32_t node_stride) {
  if (!h) return 0;
  if (node_stride < 1) node_stride = 1;
  const int G = (int)h->grid.size();
  int n = (G + node_stride - 1) / node_stride;
  if ((n - 1) * node_stride != G - 1) ++n;
  return n;
}

static int curves_on(mp_handle* h, const double* d_pars, int32_t W, int32_t ndim, int32_t node_stride, double* d_out,
                     double* d_state, int32_t* d_status, cudaStream_t stream, int lane) {
  DeviceNodes* dn = nullptr;
  int rc = curve_node_set(h, node_stride, &dn);
  if (rc) return rc;
  Problem p;
  if ((rc = fill_problem(h, p, *dn, false, ndim, W, false))) return rc;
  p.sp.unlog_mask = 0;
  return launch_eval<kModeCurves, false>(h, p, d_pars, W, Sink{nullptr, d_status, nullptr}, d_out, d_state, nullptr, stream, lane);
}

extern "C" int mp_model_curves_device(mp_handle* h, const double* d_pars, int32_t W, int32_t ndim,
                                      int32_t node_stride, double* d_out, double* d_state,
                                      int32_t* d_status, void* stream) {
  if (!h || (W > 0 && (!d_pars || !d_out))) return fail(MP_ERR_BAD_ARG, "mp_model_curves_device: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  return curves_on(h, d_pars, W, ndim, node_stride, d_out, d_state, d_status, (cudaStream_t)stream, -1);
}

extern "C" int mp_model_curves(mp_handle* h, const double* pars, int32_t W, int32_t ndim, int32_t node_stride,
                               double* out, double* state, int32_t* status) {
  if (!h || (W > 0 && (!pars || !out))) return fail(MP_ERR_BAD_ARG, "mp_model_curves: null pointer");
  if (ndim < 6 || ndim > MP_MAX_NDIM) return fail(MP_ERR_BAD_ARG, "ndim must be 6, 7, 8 or 9");
  if (W == 0) return MP_OK;
  MP_CUDA(cudaSetDevice(h->device));
  const size_t Gs = (size_t)mp_curve_nodes(h, node_stride);
  int rc;
  if ((rc = ensure(&h->s_theta, &h->cap_theta, (size_t)W * MP_MAX_NDIM))) return rc;
  if ((rc = ensure(&h->s_out, &h->cap_out, (size_t)W * 3 * Gs))) return rc;
  if (state && (rc = ensure(&h->s_state, &h->cap_state, (size_t)W * 2 * Gs))) return rc;
  int* d_status = nullptr;
  if (status) {
    if ((rc = ensure(&h->s_cstatus, &h->cap_cstatus, (size_t)W))) return rc;
    d_status = h->s_cstatus;
  }
  MP_CUDA(cudaMemcpyAsync(h->s_theta, pars, (size_t)W * ndim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  rc = curves_on(h, h->s_theta, W, ndim, node_stride, h->s_out, state ? h->s_state : nullptr, d_status, h->stream, 0);
  if (!rc) {
    cudaMemcpyAsync(out, h->s_out, (size_t)W * 3 * Gs * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (state) cudaMemcpyAsync(state, h->s_state, (size_t)W * 2 * Gs * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (status) cudaMemcpyAsync(status, d_status, (size_t)W * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(MP_ERR_CUDA, std::string("mp_model_curves: ") + cudaGetErrorString(e));
  }
  return rc;
}

extern "C" int mp_rhs_batch(const mp_model_spec* spec, const double* y, const double* t, const double* pars,
                            const double* knobs, int32_t W, double* dydt, int32_t device) {
  if (!spec || !y || !t || !pars || !knobs || !dydt || W < 0) return fail(MP_ERR_BAD_ARG, "mp_rhs_batch: null pointer");
  if (W == 0) return MP_OK;
  if (mp_device_count() <= device || device < 0)
    return fail(MP_ERR_CUDA, "mp_rhs_batch: no such CUDA device (magprop_b200 has no CPU path)");
  MP_CUDA(cudaSetDevice(device));
  // one allocation for the four arrays (y[2W] t[W] pars[5W] dydt[2W]), released on every path
  double* d_all = nullptr;
  MP_CUDA(cudaMalloc((void**)&d_all, (size_t)W * 10 * sizeof(double)));
  double *d_y = d_all, *d_t = d_all + (size_t)2 * W, *d_p = d_all + (size_t)3 * W, *d_o = d_all + (size_t)8 * W;
  cudaError_t e = cudaMemcpy(d_y, y, (size_t)W * 2 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_t, t, (size_t)W * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_p, pars, (size_t)W * 5 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rhs_kernel<<<(W + 127) / 128, 128>>>(spec->inertia_factor, spec->mdot_factor, spec->breakup_rhs, spec->dipole_torque, d_y, d_t, d_p,
                                         knobs[0], knobs[1], knobs[2], knobs[3], W, d_o);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(dydt, d_o, (size_t)W * 2 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_all);
  if (e != cudaSuccess) return fail(MP_ERR_CUDA, std::string("mp_rhs_batch: ") + cudaGetErrorString(e));
  return MP_OK;
}

static int fill_move_problem(mp_handle* h, Problem& p, int ndim, int n_active) {
  return fill_problem(h, p, h->data_nodes, true, ndim, n_active, true);
}

extern "C" int mp_stretch_half_step(mp_handle* h, double* d_coords, double* d_lnp, int32_t nwalkers,
                                    int32_t ndim, const int32_t* d_active, int32_t n_active,
                                    const int32_t* d_complement, int32_t n_complement, double a,
                                    uint64_t seed, uint64_t step, int32_t* d_accepted, int32_t* d_n_rhs,
                                    void* stream) {
  if (!h || !d_coords || !d_lnp || !d_active || !d_complement || n_active < 0 || n_complement <= 0 ||
      nwalkers <= 0)
    return fail(MP_ERR_BAD_ARG, "mp_stretch_half_step: null pointer or empty set");
  MP_CUDA(cudaSetDevice(h->device));
  Problem p;
  int rc = fill_move_problem(h, p, ndim, n_active);
  if (rc) return rc;
  Move m;
  std::memset(&m, 0, sizeof(m));
  m.coords = d_coords;
  m.lnp = d_lnp;
  m.active = d_active;
  m.complement = d_complement;
  m.n_complement = n_complement;
  m.a = a;
  m.seed = seed;
  m.step = step;
  m.accepted = d_accepted;
  m.n_rhs = d_n_rhs;
  return launch_eval<kModeLnprob, true>(h, p, nullptr, n_active, Sink{nullptr, nullptr, nullptr}, nullptr, nullptr, &m,
                                        (cudaStream_t)stream);
}

static int check_ensemble(const mp_ensemble* e, const char* who) {
  if (!e || !e->coords || !e->lnp) return fail(MP_ERR_BAD_ARG, std::string(who) + ": null ensemble / coords / lnp");
  if (e->nwalkers < 2 || (e->nwalkers & 1)) return fail(MP_ERR_BAD_ARG, std::string(who) + ": nwalkers must be even");
  if (e->world < 1 || e->rank < 0 || e->rank >= e->world || (e->nwalkers / 2) % e->world)
    return fail(MP_ERR_BAD_ARG, std::string(who) + ": half-ensemble does not split evenly over the ranks");
  if (e->n_peers < 0 || e->n_peers > MP_MAX_PEERS) return fail(MP_ERR_BAD_ARG, std::string(who) + ": bad n_peers");
  return MP_OK;
}

// The pull side of a peer-mapped ensemble: replicas, last step's order, whether the replicas are in sync.
static int fill_pull(const mp_ensemble* e, uint64_t step, Move& m, const char* who) {
  m.n_peers = e->n_peers;
  m.rank = e->rank;
  m.half = e->nwalkers / 2;
  m.n_mine = m.half / e->world;
  if (e->n_peers == 0) return MP_OK;
  if (e->n_peers != e->world - 1) return fail(MP_ERR_BAD_ARG, std::string(who) + ": n_peers must be 0 or world - 1");
  if (step < e->synced_step) return fail(MP_ERR_BAD_ARG, std::string(who) + ": step lies before synced_step");
  for (int r = 0; r < e->n_peers; ++r) {
    if (!e->peer_coords[r] || !e->peer_lnp[r]) return fail(MP_ERR_BAD_ARG, std::string(who) + ": null peer replica");
    m.peer_coords[r] = e->peer_coords[r];
    m.peer_lnp[r] = e->peer_lnp[r];
  }
  m.synced = step == e->synced_step;
  if (!m.synced) m.prev_perm = make_split_perm(e->nwalkers, e->seed, step - 1, e->randomize_split);
  return MP_OK;
}

extern "C" int mp_ensemble_sync(const mp_ensemble* e, uint64_t step, void* stream) {
  int rc = check_ensemble(e, "mp_ensemble_sync");
  if (rc) return rc;
  Move m;
  std::memset(&m, 0, sizeof(m));
  m.coords = e->coords;
  m.lnp = e->lnp;
  if ((rc = fill_pull(e, step, m, "mp_ensemble_sync"))) return rc;
  if (e->n_peers == 0 || m.synced) return MP_OK;
  sync_replica_kernel<<<(e->nwalkers + 255) / 256, 256, 0, (cudaStream_t)stream>>>(m, e->nwalkers, e->ndim);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}

extern "C" int mp_ensemble_half_step(mp_handle* h, const mp_ensemble* e, uint64_t step, int32_t split, void* stream) {
  if (!h) return fail(MP_ERR_BAD_ARG, "mp_ensemble_half_step: null handle");
  int rc = check_ensemble(e, "mp_ensemble_half_step");
  if (rc) return rc;
  if (split != 0 && split != 1) return fail(MP_ERR_BAD_ARG, "mp_ensemble_half_step: split must be 0 or 1");
  MP_CUDA(cudaSetDevice(h->device));
  const int half = e->nwalkers / 2, n_mine = half / e->world;
  Problem p;
  if ((rc = fill_move_problem(h, p, e->ndim, n_mine))) return rc;
  Move m;
  std::memset(&m, 0, sizeof(m));
  m.coords = e->coords;
  m.lnp = e->lnp;
  m.perm = make_split_perm(e->nwalkers, e->seed, step, e->randomize_split);
  m.pos0 = split * half + e->rank * n_mine;
  m.cpos0 = (1 - split) * half;
  m.n_complement = half;
  m.a = e->a;
  m.seed = e->seed;
  m.step = 2 * step + (uint64_t)split;
  m.accepted = e->accepted;
  m.status = e->status;
  m.n_rhs = e->n_rhs;
  if ((rc = fill_pull(e, step, m, "mp_ensemble_half_step"))) return rc;
  m.split = split;
  m.pack_out = e->pack_out;
  if (e->bad_rows && e->bad_count && e->bad_capacity > 0) {
    m.bad_rows = e->bad_rows;
    m.bad_count = e->bad_count;
    m.bad_capacity = e->bad_capacity;
  }
  return launch_eval<kModeLnprob, true>(h, p, nullptr, n_mine, Sink{nullptr, nullptr, nullptr}, nullptr, nullptr, &m,
                                        (cudaStream_t)stream);
}

extern "C" int mp_ensemble_unpack(const mp_ensemble* e, uint64_t step, int32_t split, const double* d_packed, void* stream) {
  int rc = check_ensemble(e, "mp_ensemble_unpack");
  if (rc) return rc;
  if (!d_packed || (split != 0 && split != 1)) return fail(MP_ERR_BAD_ARG, "mp_ensemble_unpack: bad argument");
  const int half = e->nwalkers / 2;
  const SplitPerm perm = make_split_perm(e->nwalkers, e->seed, step, e->randomize_split);
  unpack_kernel<<<(half + 255) / 256, 256, 0, (cudaStream_t)stream>>>(perm, split * half, half, e->ndim, d_packed, e->coords, e->lnp);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}

extern "C" int mp_ensemble_order(int32_t nwalkers, uint64_t seed, uint64_t step, int32_t randomize_split,
                                 int32_t* d_order, void* stream) {
  if (nwalkers <= 0 || !d_order) return fail(MP_ERR_BAD_ARG, "mp_ensemble_order: bad argument");
  const SplitPerm perm = make_split_perm(nwalkers, seed, step, randomize_split);
  order_kernel<<<(nwalkers + 255) / 256, 256, 0, (cudaStream_t)stream>>>(perm, nwalkers, d_order);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}

// ---- peer-mapped replicas (CUDA IPC) -------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI carries IPC handles as 64 bytes");

extern "C" int mp_peer_alloc(int32_t device, uint64_t bytes, void** d_ptr, unsigned char handle[64]) {
  if (!d_ptr || !handle || bytes == 0) return fail(MP_ERR_BAD_ARG, "mp_peer_alloc: bad argument");
  if (mp_device_count() <= device || device < 0) return fail(MP_ERR_CUDA, "mp_peer_alloc: no such CUDA device");
  MP_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  MP_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t hd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(MP_ERR_CUDA, std::string("mp_peer_alloc: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle, &hd, 64);
  *d_ptr = p;
  return MP_OK;
}

extern "C" int mp_peer_open(int32_t device, const unsigned char handle[64], void** d_ptr) {
  if (!d_ptr || !handle) return fail(MP_ERR_BAD_ARG, "mp_peer_open: null pointer");
  MP_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, 64);
  MP_CUDA(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return MP_OK;
}

extern "C" int mp_peer_close(int32_t device, void* d_ptr) {
  if (!d_ptr) return MP_OK;
  MP_CUDA(cudaSetDevice(device));
  MP_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return MP_OK;
}

extern "C" int mp_peer_free(int32_t device, void* d_ptr) {
  if (!d_ptr) return MP_OK;
  MP_CUDA(cudaSetDevice(device));
  MP_CUDA(cudaFree(d_ptr));
  return MP_OK;
}

extern "C" int mp_peer_barrier(int32_t device, uint64_t* d_my_flags, uint64_t* const* peer_flags, int32_t rank,
                               int32_t world, uint64_t epoch, int32_t* d_error, void* stream) {
  if (!d_my_flags || !peer_flags || world < 1 || world > MP_MAX_PEERS + 1 || rank < 0 || rank >= world)
    return fail(MP_ERR_BAD_ARG, "mp_peer_barrier: bad argument");
  if (world == 1) return MP_OK;
  MP_CUDA(cudaSetDevice(device));
  PeerFlags pf;
  std::memset(&pf, 0, sizeof(pf));
  for (int p = 0; p < world; ++p) {
    if (p != rank && !peer_flags[p]) return fail(MP_ERR_BAD_ARG, "mp_peer_barrier: null peer flag array");
    pf.p[p] = peer_flags[p];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_my_flags, pf, rank, world, epoch, d_error);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}

// ---- posterior summaries on the device ---------------------------------------------------------------------
extern "C" int mp_chain_moments(const double* d_chain, int64_t n, int32_t ndim, double* mean, double* cov, int32_t device,
                                void* stream) {
  if (!d_chain || !mean || !cov || n < 1 || ndim < 1 || ndim > MP_MAX_NDIM) return fail(MP_ERR_BAD_ARG, "mp_chain_moments: bad argument");
  if (mp_device_count() <= device || device < 0) return fail(MP_ERR_CUDA, "mp_chain_moments: no such CUDA device");
  MP_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  double* d_acc = nullptr;
  MP_CUDA(cudaMalloc((void**)&d_acc, (ndim + ndim * ndim) * sizeof(double)));
  cudaError_t e = cudaMemsetAsync(d_acc, 0, (ndim + ndim * ndim) * sizeof(double), st);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 1184);
  std::vector<double> hm(ndim), hc((size_t)ndim * ndim);
  if (e == cudaSuccess) {
    chain_mean_kernel<<<blocks, 256, 0, st>>>(d_chain, n, ndim, d_acc);
    e = cudaMemcpyAsync(hm.data(), d_acc, ndim * sizeof(double), cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) {
    for (int d = 0; d < ndim; ++d) hm[d] /= (double)n;
    e = cudaMemcpyAsync(d_acc, hm.data(), ndim * sizeof(double), cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess) {
    chain_cov_kernel<<<blocks, 256, 0, st>>>(d_chain, n, ndim, d_acc, d_acc + ndim);
    e = cudaMemcpyAsync(hc.data(), d_acc + ndim, (size_t)ndim * ndim * sizeof(double), cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_acc);
  if (e != cudaSuccess) return fail(MP_ERR_CUDA, std::string("mp_chain_moments: ") + cudaGetErrorString(e));
  for (int a = 0; a < ndim; ++a) {
    mean[a] = hm[a];
    for (int b = a; b < ndim; ++b) cov[a * ndim + b] = cov[b * ndim + a] = hc[(size_t)a * ndim + b] / (double)(n - 1 > 0 ? n - 1 : 1);
  }
  return MP_OK;
}

extern "C" int mp_chain_order_statistics(const double* d_chain, int64_t n, int32_t ndim, int32_t col, const int64_t* ranks,
                                         int32_t n_ranks, double* values, int32_t device, void* stream) {
  if (!d_chain || !ranks || !values || n < 1 || ndim < 1 || col < 0 || col >= ndim || n_ranks < 0)
    return fail(MP_ERR_BAD_ARG, "mp_chain_order_statistics: bad argument");
  if (mp_device_count() <= device || device < 0) return fail(MP_ERR_CUDA, "mp_chain_order_statistics: no such CUDA device");
  MP_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* d_hist = nullptr;
  MP_CUDA(cudaMalloc((void**)&d_hist, 65536 * sizeof(unsigned)));
  std::vector<unsigned> hist(65536);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 1184);
  cudaError_t e = cudaSuccess;
  for (int r = 0; r < n_ranks && e == cudaSuccess; ++r) {
    int64_t k = ranks[r];
    if (k < 0 || k >= n) { cudaFree(d_hist); return fail(MP_ERR_BAD_ARG, "mp_chain_order_statistics: rank outside [0, n)"); }
    unsigned long long prefix = 0ull;
    for (int shift = 48; shift >= 0 && e == cudaSuccess; shift -= 16) {
      e = cudaMemsetAsync(d_hist, 0, 65536 * sizeof(unsigned), st);
      if (e != cudaSuccess) break;
      select_hist_kernel<<<blocks, 256, 0, st>>>(d_chain, n, ndim, col, prefix, shift, d_hist);
      e = cudaMemcpyAsync(hist.data(), d_hist, 65536 * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) break;
      unsigned digit = 0;
      for (; digit < 65536u; ++digit) {
        if (k < (int64_t)hist[digit]) break;
        k -= hist[digit];
      }
      prefix |= (unsigned long long)digit << shift;
    }
    // key -> double
    const unsigned long long b = (prefix >> 63) ? (prefix & 0x7fffffffffffffffull) : ~prefix;
    std::memcpy(values + r, &b, sizeof(double));
  }
  cudaFree(d_hist);
  if (e != cudaSuccess) return fail(MP_ERR_CUDA, std::string("mp_chain_order_statistics: ") + cudaGetErrorString(e));
  return MP_OK;
}

extern "C" int mp_gompertz_curves(const double* pars, int32_t W, const double* knobs, int64_t n_steps, int32_t stride,
                                  double* out, int32_t device) {
  if (!pars || !knobs || !out || W < 0 || n_steps < 1 || stride < 1) return fail(MP_ERR_BAD_ARG, "mp_gompertz_curves: bad argument");
  if (W == 0) return MP_OK;
  if (mp_device_count() <= device || device < 0)
    return fail(MP_ERR_CUDA, "mp_gompertz_curves: no such CUDA device (magprop_b200 has no CPU path)");
  MP_CUDA(cudaSetDevice(device));
  const long long n_out = (n_steps + stride - 1) / stride;
  double *d_p = nullptr, *d_o = nullptr;
  MP_CUDA(cudaMalloc((void**)&d_p, (size_t)W * 6 * sizeof(double)));
  cudaError_t e = cudaMalloc((void**)&d_o, (size_t)W * 3 * n_out * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d_p, pars, (size_t)W * 6 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    gompertz_kernel<<<(W + 31) / 32, 32>>>(d_p, W, knobs[0], knobs[1], knobs[2], knobs[3], knobs[4], knobs[5], n_steps, stride, n_out, d_o);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_o, (size_t)W * 3 * n_out * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_p);
  cudaFree(d_o);
  if (e != cudaSuccess) return fail(MP_ERR_CUDA, std::string("mp_gompertz_curves: ") + cudaGetErrorString(e));
  return MP_OK;
}

extern "C" int mp_last_stiff_count(mp_handle* h, int32_t* count) {
  if (!h || !count) return fail(MP_ERR_BAD_ARG, "mp_last_stiff_count: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  MP_CUDA(cudaDeviceSynchronize());
  *count = 0;
  if (h->last_lane && h->last_lane->counters)
    MP_CUDA(cudaMemcpy(count, h->last_lane->counters + 2, sizeof(int), cudaMemcpyDeviceToHost));
  return MP_OK;
}

extern "C" int mp_fp64_peak_tflops(int32_t device, double* tflops) {
  if (!tflops) return fail(MP_ERR_BAD_ARG, "mp_fp64_peak_tflops: null pointer");
  if (mp_device_count() <= device || device < 0) return fail(MP_ERR_CUDA, "no such CUDA device");
  MP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MP_CUDA(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  MP_CUDA(cudaMalloc((void**)&d, 8));
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  cudaEvent_t e0, e1;
  MP_CUDA(cudaEventCreate(&e0));
  MP_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    MP_CUDA(cudaEventRecord(e0));
    dfma_peak_kernel<<<blocks, 256>>>(d, iters, 1.0 + rep);
    MP_CUDA(cudaEventRecord(e1));
    MP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return MP_OK;
}
