// magprop_kernels.cu -- sm_100a kernels and the C ABI of include/magprop_b200.h.
//
// Kernels (one walker per thread, FP64 throughout, no tensor cores -- the path
// is a scalar ODE solve, not a contraction):
//   eval_kernel<kModeLnprob>      prior -> spin ODE -> luminosity -> interp -> chi2 -> lnprob
//   eval_kernel<kModeModelAtData> same, writes the model at the data times
//   eval_kernel<kModeCurves>      same, writes Ltot/Lprop/Ldip (and state) at grid nodes
//   stretch_kernel                emcee stretch-move proposal + the above + accept, fused
//   rhs_kernel                    the coupled reference RHS for ODEs()/odes() callers
//   dfma_peak_kernel              FP64 FMA roofline denominator
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "magprop_host.hpp"
#include "magprop_rng.cuh"

namespace mp {

#ifndef MP_KNB
#define MP_KNB 16
#endif
constexpr int kNB = MP_KNB;  // nodes buffered per thread between phase A and phase B
constexpr int kResumeLen = kResumeDoubles + kNB;   // one hand-over record (see evaluate_walker)
// curve output hands one node to each lane of the warp in phase B, so it buffers a full warp's worth
template <int MODE> struct NodeBuf { static constexpr int n = (MODE == kModeCurves) ? 32 : kNB; };

struct KernelArgs {
  Spec sp;
  DataView dv;
  const int* dat_orig;  // sorted datum -> caller's index (kModeModelAtData)
  double lower[MP_MAX_NDIM], upper[MP_MAX_NDIM];
  int prior_enabled;
  int ndim;
  int W;
  const double* theta;  // [W][ndim]
  double* lnp;          // [W]
  int* status;          // [W] or null
  int* n_rhs;           // [W] or null
  double* out;          // mode-dependent
  double* state;        // curves: [W][2][Gs] or null
  const int* order;     // [W] or null: thread i evaluates walker order[i] (walkers bucketed by a cost key)
  int lanes_per_walker; // curve output: one walker per this many lanes (1, 2, .. 32), see launch_eval
  int* queue;           // [W] walkers deferred to the stiff launch
  int* queue_count;     // [1]
  double* resume;       // [queue capacity][kResumeLen] hand-over records of the deferred walkers, or null
};

// Coalesced load of this block's walker parameters into shared memory
// (theta rows are 48..72 B apart, so per-thread row reads would be 8-byte
// gathers).
template <int BLOCK>
__device__ __forceinline__ void stage_theta(const double* __restrict__ theta, int W, int ndim,
                                            double* s_theta, const int* __restrict__ order, int lpw) {
  if (order || lpw > 1) {      // bucketed launch / sparse warps: each thread gathers its own row
    const int slot = blockIdx.x * BLOCK + threadIdx.x;
    const int i = slot / lpw;
    const bool have = (i < W) && (slot % lpw == 0);
    const long long w = have ? (order ? order[i] : i) : 0;
    for (int d = 0; d < ndim; ++d) s_theta[threadIdx.x * ndim + d] = have ? theta[w * ndim + d] : 0.0;
    __syncthreads();
    return;
  }
  const long long base = (long long)blockIdx.x * BLOCK * ndim;
  const long long total = (long long)W * ndim;
  for (int i = threadIdx.x; i < BLOCK * ndim; i += BLOCK) {
    const long long g = base + i;
    s_theta[i] = (g < total) ? theta[g] : 0.0;
  }
  __syncthreads();
}

// The dataset a block works against -- node times and, per datum, y/yerr, 1e-50/yerr, x - t_lo, the
// interpolation weight and the lower-node index -- staged in shared memory when it fits the budget below
// (the synthetic datasets take 2.2 KB; a 1944-point burst stays in global memory behind L1).
constexpr int kDataSmemDoubles = 768;      // 6 KB per block
__device__ __forceinline__ bool stage_data(const DataView& g, DataView& s, double* buf, int nthreads) {
  const int Nn = g.n_nodes, D = g.n_data;
  const int need = Nn + 4 * D + (D + 1) / 2;
  if (need > kDataSmemDoubles) return false;
  double* p = buf;
  double* nt = p; p += Nn;
  double* ys = p; p += D;
  double* c = p; p += D;
  double* dx = p; p += D;
  double* w = p; p += D;
  int* lo = reinterpret_cast<int*>(p);
  for (int i = threadIdx.x; i < Nn; i += nthreads) nt[i] = g.node_t[i];
  for (int i = threadIdx.x; i < D; i += nthreads) {
    ys[i] = g.dat_ys[i]; c[i] = g.dat_c[i]; dx[i] = g.dat_dx[i]; w[i] = g.dat_w[i]; lo[i] = g.dat_lo[i];
  }
  s = g;
  s.node_t = nt; s.dat_ys = ys; s.dat_c = c; s.dat_dx = dx; s.dat_w = w; s.dat_lo = lo;
  return true;
}

#ifndef MP_MIN_BLOCKS_32
#define MP_MIN_BLOCKS_32 16
#endif
#ifndef MP_MIN_BLOCKS_64
#define MP_MIN_BLOCKS_64 8
#endif

// One walker, end to end.  Returns true when the walker was deferred to the stiff launch.
// `have` = false: this lane has no walker; it still walks through evaluate_walker as a bystander
// because the warp votes there need every lane.
template <int MODE, int BLOCK, bool STIFF>
__device__ __forceinline__ bool eval_one(const KernelArgs& a, const DataView& dv, bool have, int w, const double* th, double* s_buf,
                                         void* warp_scratch, ResumeSink* sink = nullptr, const double* rec_in = nullptr) {
  int st = kWalkerOk, nr = 0;
  double result = -INFINITY;
  const bool rejected = have && a.prior_enabled && !prior_accepts(th, a.ndim, a.lower, a.upper);
  if (rejected) st = kWalkerPriorReject;           // mcmc_eqns.py:66-69: model is skipped
  const bool live = have && !rejected;
  {
    double pars[6], dipeff, propeff, f_beam;
    unpack_theta(a.sp, th, a.ndim, pars, dipeff, propeff, f_beam);
    Walker wk;
    walker_setup(a.sp, pars, dipeff, propeff, f_beam, dv.t_start, wk);
    double* out = nullptr;
    double* state = nullptr;
    if (MODE == kModeCurves) {
      out = a.out + (size_t)w * 3 * dv.n_nodes;
      state = a.state ? a.state + (size_t)w * 2 * dv.n_nodes : nullptr;
    } else if (MODE == kModeModelAtData) {
      out = a.out + (size_t)w * dv.n_data;
    }
    const double chi2 = evaluate_walker<MODE, NodeBuf<MODE>::n, STIFF>(a.sp, dv, wk, live, s_buf + threadIdx.x, BLOCK, st, nr,
                                                          out, state, 1, a.dat_orig, warp_scratch,
                                                          MODE == kModeCurves ? nullptr : sink,
                                                          MODE == kModeCurves ? nullptr : rec_in);
    if (!STIFF && (st & kWalkerDeferred)) return true;
    if (live && MODE == kModeLnprob) {
      double ll = -0.5 * chi2;                     // mcmc_eqns.py:25
      if (st & kWalkerIntegratorFail) {
        ll = -INFINITY;                            // 'flag' -> -inf (mcmc_eqns.py:22-23)
      } else if (!isfinite(ll)) {
        st |= kWalkerNonfiniteLnlike;              // mcmc_eqns.py:72-79
        ll = -INFINITY;
      }
      result = ll;                                 // + lnprior == 0.0
    }
  }
  if (!have) return false;
  if (MODE == kModeLnprob) a.lnp[w] = result;
  if (a.status) a.status[w] = st;
  if (a.n_rhs) a.n_rhs[w] = nr;
  return false;
}

// Main launch: explicit integrator only; stiff walkers are pushed onto a queue.
template <int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 32 ? MP_MIN_BLOCKS_32 : MP_MIN_BLOCKS_64))
eval_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ double s_buf[NodeBuf<MODE>::n * BLOCK];
  __shared__ double s_theta[BLOCK * MP_MAX_NDIM];
  __shared__ Walker s_walker[MODE == kModeCurves ? BLOCK / 32 : 1];   // per-warp broadcast slot (curve output)
  __shared__ double s_data[MODE == kModeCurves ? 1 : kDataSmemDoubles];
  DataView dv = a.dv;
  if (MODE != kModeCurves) stage_data(a.dv, dv, s_data, BLOCK);       // (stage_theta's barrier covers it)
  const int lpw = (MODE == kModeCurves && a.lanes_per_walker > 1) ? a.lanes_per_walker : 1;
  stage_theta<BLOCK>(a.theta, a.W, a.ndim, s_theta, a.order, lpw);
  const int slot = blockIdx.x * BLOCK + threadIdx.x;
  const int i = slot / lpw;
  const bool have = (i < a.W) && (slot % lpw == 0);
  const int w = (a.order && have) ? a.order[i] : i;
  void* scratch = &s_walker[MODE == kModeCurves ? (threadIdx.x >> 5) : 0];
  // (a deferred walker claims its queue slot -- and leaves its hand-over record there -- inside evaluate_walker)
  ResumeSink sink{(MODE != kModeCurves && a.resume) ? a.queue_count : nullptr, a.resume, -1};
  if (eval_one<MODE, BLOCK, false>(a, dv, have, w, s_theta + threadIdx.x * a.ndim, s_buf, scratch, &sink))
    a.queue[sink.count ? sink.slot : atomicAdd(a.queue_count, 1)] = w;
}

// Second launch: the walkers bucketed as stiff, with the implicit integrator available.
#ifndef MP_STIFF_MIN_WARPS
#define MP_STIFF_MIN_WARPS 12     // resident warps per SM the implicit variant is compiled for
#endif
template <int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MP_STIFF_MIN_WARPS * 32 / BLOCK) eval_stiff_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ double s_buf[NodeBuf<MODE>::n * BLOCK];
  __shared__ Walker s_walker[MODE == kModeCurves ? BLOCK / 32 : 1];
  void* scratch = &s_walker[MODE == kModeCurves ? (threadIdx.x >> 5) : 0];
  // One batch per block and a grid sized for the whole launch (the queue length is only known on the
  // device): blocks beyond the queue exit at once, and the hardware block scheduler balances the
  // very unequal walkers of this bucket.
  const int n = *a.queue_count;
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (blockIdx.x * BLOCK >= n) return;
  const bool have = i < n;
  const int w = have ? a.queue[i] : 0;
  double th[MP_MAX_NDIM];
  for (int d = 0; d < a.ndim; ++d) th[d] = a.theta[(size_t)w * a.ndim + d];
  eval_one<MODE, BLOCK, true>(a, a.dv, have, w, th, s_buf, scratch, nullptr,
                              (MODE != kModeCurves && a.resume && have) ? a.resume + (size_t)i * kResumeLen : nullptr);
}

struct StretchArgs {
  KernelArgs k;          // k.theta unused; k.lnp unused
  double* coords;        // [nwalkers][ndim], updated in place
  double* lnp;           // [nwalkers]
  // who moves against whom: explicit index lists (mp_stretch_half_step) ...
  const int* active;     // [n_active] walkers to move (disjoint from complement), or null
  const int* complement; // [n_complement]
  // ... or positions of the ensemble order (mp_ensemble_half_step): mover i is walker P(pos0 + i), its
  // partner is drawn from P(cpos0 + [0, n_complement))
  SplitPerm perm;
  int pos0, cpos0;
  int n_active, n_complement;
  double a;
  uint64_t seed, step;   // step: the half-step counter (RNG counter word)
  int* accepted;         // [nwalkers] counters (may be null)
  int* status;           // [nwalkers] MP_WALKER_* bits of the latest proposal (may be null)
  // replicas of (coords, lnp) on the other ranks, peer-mapped over NVLink: accepted rows are stored there too
  int n_peers;
  double* peer_coords[MP_MAX_PEERS];
  double* peer_lnp[MP_MAX_PEERS];
  double* pack_out;      // [n_active][ndim+1]: every mover's (row, lnp) after the move, or null
  double* bad_rows;      // proposals whose likelihood was not finite ({GRB}_bad.csv, mcmc_eqns.py:72-79)
  int* bad_count;
  int bad_capacity;
};

// One emcee StretchMove half-step (Goodman & Weare 2010; emcee RedBlueMove):
//   z = ((a-1) u + 1)^2 / a ; q = c - (c - s) z ; accept iff (ndim-1) ln z + lp(q) - lp(s) > ln u'
// fused with the likelihood so a half-step is one launch (plus the stiff-bucket launch, which
// finds an empty queue for ensembles near the synthetic truths).  `i` is the mover's index in this
// launch.  Returns true when deferred.
template <int BLOCK, bool STIFF>
__device__ __forceinline__ bool stretch_one(const StretchArgs& s, bool have, int i, double* s_buf, ResumeSink* sink = nullptr,
                                            const double* rec_in = nullptr) {
  const KernelArgs& a = s.k;
  const int ndim = a.ndim;
  const int me = s.active ? s.active[i] : (int)perm_at(s.perm, (uint32_t)(s.pos0 + i));
  // counter = (half-step, walker); two Philox blocks give u_z, u_partner, u_accept
  const Philox r0 = philox4x32_10((uint32_t)s.step, (uint32_t)(s.step >> 32), (uint32_t)me, 0u,
                                  (uint32_t)s.seed, (uint32_t)(s.seed >> 32));
  const Philox r1 = philox4x32_10((uint32_t)s.step, (uint32_t)(s.step >> 32), (uint32_t)me, 1u,
                                  (uint32_t)s.seed, (uint32_t)(s.seed >> 32));
  const double uz = u01(r0.c[0], r0.c[1]);
  const double up = u01(r0.c[2], r0.c[3]);
  const double ua = u01(r1.c[0], r1.c[1]);
  const double zr = __dadd_rn(__dmul_rn(s.a - 1.0, uz), 1.0);
  const double z = __ddiv_rn(__dmul_rn(zr, zr), s.a);
  int pj = (int)(up * s.n_complement);
  if (pj >= s.n_complement) pj = s.n_complement - 1;
  const int partner = s.complement ? s.complement[pj] : (int)perm_at(s.perm, (uint32_t)(s.cpos0 + pj));
  double q[MP_MAX_NDIM];
  for (int d = 0; d < ndim; ++d) {
    const double c = s.coords[(size_t)partner * ndim + d];
    const double x = s.coords[(size_t)me * ndim + d];
    q[d] = __dadd_rn(c, -__dmul_rn(__dadd_rn(c, -x), z));   // no FMA contraction: reproducible on the host
  }
  int st = kWalkerOk, nr = 0;
  double lp_new = -INFINITY;
  const bool live = have && (!a.prior_enabled || prior_accepts(q, ndim, a.lower, a.upper));
  if (have && !live) st = kWalkerPriorReject;
  {
    double pars[6], dipeff, propeff, f_beam;
    unpack_theta(a.sp, q, ndim, pars, dipeff, propeff, f_beam);
    Walker wk;
    walker_setup(a.sp, pars, dipeff, propeff, f_beam, a.dv.t_start, wk);
    const double chi2 = evaluate_walker<kModeLnprob, kNB, STIFF>(a.sp, a.dv, wk, live, s_buf + threadIdx.x, BLOCK,
                                                                 st, nr, nullptr, nullptr, 1, nullptr, nullptr, sink, rec_in);
    if (!STIFF && (st & kWalkerDeferred)) return true;
    if (live) {
      double ll = -0.5 * chi2;
      if (st & kWalkerIntegratorFail) {
        ll = -INFINITY;
      } else if (!isfinite(ll)) {
        st |= kWalkerNonfiniteLnlike;
        ll = -INFINITY;
      }
      lp_new = ll;
    }
  }
  if (!have) return false;
  const double lp_old = s.lnp[me];
  const double lnpdiff = __dadd_rn(__dadd_rn(__dmul_rn(ndim - 1.0, log(z)), lp_new), -lp_old);
  const bool accept = lnpdiff > log(ua);
  if (accept) {
    for (int d = 0; d < ndim; ++d) s.coords[(size_t)me * ndim + d] = q[d];
    s.lnp[me] = lp_new;
    // the same row into every other rank's replica (NVLink peer stores; the caller's mp_peer_barrier
    // makes them visible before the next half-step reads them)
    for (int p = 0; p < s.n_peers; ++p) {
      double* pc = s.peer_coords[p] + (size_t)me * ndim;
      for (int d = 0; d < ndim; ++d) pc[d] = q[d];
      s.peer_lnp[p][me] = lp_new;
    }
    if (s.accepted) s.accepted[me] += 1;
  }
  if (s.pack_out) {
    double* row = s.pack_out + (size_t)i * (ndim + 1);
    for (int d = 0; d < ndim; ++d) row[d] = accept ? q[d] : s.coords[(size_t)me * ndim + d];
    row[ndim] = accept ? lp_new : lp_old;
  }
  if (s.status) s.status[me] = st;
  if (s.bad_count && (st & (kWalkerIntegratorFail | kWalkerNonfiniteLnlike))) {
    const int slot = atomicAdd(s.bad_count, 1);
    if (slot < s.bad_capacity)
      for (int d = 0; d < ndim; ++d) s.bad_rows[(size_t)slot * ndim + d] = q[d];
  }
  if (a.n_rhs) a.n_rhs[me] = nr;
  return false;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 32 ? MP_MIN_BLOCKS_32 : MP_MIN_BLOCKS_64))
stretch_kernel(const __grid_constant__ StretchArgs s) {
  __shared__ double s_buf[kNB * BLOCK];
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  const bool have = i < s.n_active;
  ResumeSink sink{s.k.resume ? s.k.queue_count : nullptr, s.k.resume, -1};
  if (stretch_one<BLOCK, false>(s, have, have ? i : 0, s_buf, &sink))
    s.k.queue[sink.count ? sink.slot : atomicAdd(s.k.queue_count, 1)] = i;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, MP_STIFF_MIN_WARPS * 32 / BLOCK) stretch_stiff_kernel(const __grid_constant__ StretchArgs s) {
  __shared__ double s_buf[kNB * BLOCK];
  const int n = *s.k.queue_count;
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (blockIdx.x * BLOCK >= n) return;
  const bool have = i < n;
  stretch_one<BLOCK, true>(s, have, have ? s.k.queue[i] : 0, s_buf, nullptr,
                           (s.k.resume && have) ? s.k.resume + (size_t)i * kResumeLen : nullptr);
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
  // Two pipeline lanes: each has its own stream and its own stiff-walker queue, so that the
  // host-pointer entry points can overlap the H2D copy of one chunk with the kernel of the previous
  // one.  Device-pointer entry points use lane 0's queue on the caller's stream.
  struct Lane {
    cudaStream_t stream = nullptr;
    int* queue = nullptr;        // walkers deferred to the stiff launch
    int* queue_count = nullptr;
    size_t cap_queue = 0;
    double* resume = nullptr;    // hand-over records, one per queue slot
    size_t cap_resume = 0;
    // walker bucketing (mp_set_bucketing): sort keys / walker ids (double-buffered) and CUB's scratch
    unsigned *key_in = nullptr, *key_out = nullptr;
    int *id_in = nullptr, *id_out = nullptr;
    void* sort_tmp = nullptr;
    size_t cap_sort = 0, cap_tmp = 0;
  } lanes[2];
  // Device-pointer entry points run on the caller's stream.  Each stream gets its own queue set, so calls
  // on one handle from different streams (several ensembles on one dataset, a device call next to an
  // mp_lnprob_batch_async in flight) never share a stiff queue; calls on ONE stream are ordered by the stream.
  std::map<cudaStream_t, Lane> user_lanes;
  std::mutex user_lanes_mu;
  int bucketing = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;   // == lanes[0].stream
  Lane* last_lane = nullptr;       // queue set of the most recent device-pointer launch (mp_last_stiff_count)
};

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return MP_OK;
  MP_CUDA(cudaMalloc((void**)dst, v.size() * sizeof(T)));
  MP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MP_OK;
}

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
    if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&L.queue_count, sizeof(int)) != cudaSuccess ||
        cudaMemset(L.queue_count, 0, sizeof(int)) != cudaSuccess) {
      mp_destroy(h);
      return fail(MP_ERR_CUDA, "mp_create: stream / queue allocation failed");
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
  cudaFree(h->data_nodes.node_t);
  cudaFree(h->d_y); cudaFree(h->d_yerr); cudaFree(h->d_dx); cudaFree(h->d_Dx);
  cudaFree(h->d_lo); cudaFree(h->d_orig);
  for (auto& kv : h->curve_nodes) cudaFree(kv.second.node_t);
  cudaFree(h->s_theta); cudaFree(h->s_out); cudaFree(h->s_state); cudaFree(h->s_lnp);
  cudaFree(h->s_status); cudaFree(h->s_nrhs); cudaFree(h->s_cstatus);
  auto free_lane = [](mp_handle::Lane& L) {
    cudaFree(L.queue); cudaFree(L.queue_count); cudaFree(L.resume);
    cudaFree(L.key_in); cudaFree(L.key_out); cudaFree(L.id_in); cudaFree(L.id_out); cudaFree(L.sort_tmp);
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

extern "C" int mp_set_bucketing(mp_handle* h, int32_t enabled) {
  if (!h) return fail(MP_ERR_BAD_ARG, "mp_set_bucketing: null handle");
  h->bucketing = enabled ? 1 : 0;
  return MP_OK;
}

static int fill_args(mp_handle* h, KernelArgs& a, const DeviceNodes& nodes, bool with_data, int ndim, int W,
                     bool use_prior) {
  if (ndim < 6 || ndim > MP_MAX_NDIM) return fail(MP_ERR_BAD_ARG, "ndim must be 6, 7, 8 or 9");
  if (W < 0) return fail(MP_ERR_BAD_ARG, "negative walker count");
  if (use_prior && h->prior.enabled && h->prior.ndim != ndim)
    return fail(MP_ERR_BAD_ARG, "prior dimension does not match ndim");
  std::memset(&a, 0, sizeof(a));
  a.sp = h->spec;
  a.dv.n_nodes = nodes.n_nodes;
  a.dv.node_t = nodes.node_t;
  a.dv.t_start = h->grid[0];
  if (with_data) {
    a.dv.n_data = h->D;
    a.dv.dat_ys = h->d_y;
    a.dv.dat_c = h->d_yerr;
    a.dv.dat_dx = h->d_dx;
    a.dv.dat_w = h->d_Dx;
    a.dv.dat_lo = h->d_lo;
    a.dat_orig = h->d_orig;
  }
  a.prior_enabled = use_prior ? h->prior.enabled : 0;
  for (int i = 0; i < MP_MAX_NDIM; ++i) {
    a.lower[i] = h->prior.lower[i];
    a.upper[i] = h->prior.upper[i];
  }
  a.ndim = ndim;
  a.W = W;
  return MP_OK;
}

// lane >= 0: one of the handle's own pipeline lanes; lane < 0: the queue set of the caller's stream
static int lane_for(mp_handle* h, cudaStream_t stream, int lane, mp_handle::Lane** out) {
  if (lane >= 0) {
    *out = &h->lanes[lane];
    return MP_OK;
  }
  std::lock_guard<std::mutex> g(h->user_lanes_mu);
  mp_handle::Lane& L = h->user_lanes[stream];
  if (!L.queue_count) {
    MP_CUDA(cudaMalloc((void**)&L.queue_count, sizeof(int)));
    MP_CUDA(cudaMemsetAsync(L.queue_count, 0, sizeof(int), stream));
  }
  h->last_lane = &L;
  *out = &L;
  return MP_OK;
}

static int prepare_queue(mp_handle* h, KernelArgs& a, int W, cudaStream_t stream, int lane = -1) {
  mp_handle::Lane* Lp = nullptr;
  int rc = lane_for(h, stream, lane, &Lp);
  if (rc) return rc;
  mp_handle::Lane& L = *Lp;
  rc = ensure(&L.queue, &L.cap_queue, (size_t)W);
  if (rc) return rc;
  if ((rc = ensure(&L.resume, &L.cap_resume, (size_t)W * kResumeLen))) return rc;
  a.queue = L.queue;
  a.queue_count = L.queue_count;
  a.resume = L.resume;
  MP_CUDA(cudaMemsetAsync(L.queue_count, 0, sizeof(int), stream));
  return MP_OK;
}

// grid of the stiff-bucket launch: one block per batch of the longest possible queue
static int stiff_grid(const mp_handle*, int W, int block) { return (W + block - 1) / block; }

// ---- walker bucketing ------------------------------------------------------------------------
// The step count of a walker grows with the mass that flows through the disc (log MdiscI + log delta:
// correlation 0.58 / 0.52 with log-steps over the prior box) and, second, with the disc
// radius (the viscous time).  In an ensemble that is spread out the lanes of a warp therefore finish at
// very different times (prior-uniform: the slowest lane of a warp does 3.2x the mean).  With bucketing on,
// the walkers of a launch are ordered by that key (quarter-decade bins of the mass flow, then radius) and
// thread i evaluates walker order[i]: similar walkers share a warp.  Measured on B200: prior-uniform
// ensembles +35 %, posterior-like spreads +5..20 %, a 1e-4 ball -2 % (the sort) -- hence opt-in.
__global__ void bucket_key_kernel(const double* __restrict__ theta, int W, int ndim, int unlog_mask,
                                  unsigned* __restrict__ key, int* __restrict__ id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const double* th = theta + (size_t)i * ndim;
  const double lm = ((unlog_mask >> 2) & 1) ? th[2] : log10(th[2]);
  const double lr = ((unlog_mask >> 3) & 1) ? th[3] : log10(th[3]);
  const double ld = ((unlog_mask >> 5) & 1) ? th[5] : log10(th[5]);
  const double flow = fmin(fmax((lm + ld + 16.0) * 4.0, 0.0), 255.0);     // quarter-decade bins
  const double rad = fmin(fmax(lr * 8192.0, 0.0), 65535.0);
  const unsigned k = ((unsigned)(flow == flow ? flow : 0.0) << 16) | (unsigned)(rad == rad ? rad : 0.0);
  key[i] = k;
  id[i] = i;
}

static int bucket_walkers(mp_handle* h, KernelArgs& a, cudaStream_t stream, int lane) {
  mp_handle::Lane* Lp = nullptr;
  int rc0 = lane_for(h, stream, lane, &Lp);
  if (rc0) return rc0;
  mp_handle::Lane& L = *Lp;
  const size_t W = (size_t)a.W;
  if (W > L.cap_sort) {
    cudaFree(L.key_in); cudaFree(L.key_out); cudaFree(L.id_in); cudaFree(L.id_out);
    L.key_in = L.key_out = nullptr; L.id_in = L.id_out = nullptr; L.cap_sort = 0;
    MP_CUDA(cudaMalloc((void**)&L.key_in, W * sizeof(unsigned)));
    MP_CUDA(cudaMalloc((void**)&L.key_out, W * sizeof(unsigned)));
    MP_CUDA(cudaMalloc((void**)&L.id_in, W * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&L.id_out, W * sizeof(int)));
    L.cap_sort = W;
  }
  size_t need = 0;
  MP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, L.key_in, L.key_out, L.id_in, L.id_out, (int)W, 0, 24, stream));
  if (need > L.cap_tmp) {
    cudaFree(L.sort_tmp);
    L.sort_tmp = nullptr; L.cap_tmp = 0;
    MP_CUDA(cudaMalloc(&L.sort_tmp, need));
    L.cap_tmp = need;
  }
  bucket_key_kernel<<<(a.W + 255) / 256, 256, 0, stream>>>(a.theta, a.W, a.ndim, a.sp.unlog_mask, L.key_in, L.id_in);
  MP_CUDA(cub::DeviceRadixSort::SortPairs(L.sort_tmp, need, L.key_in, L.key_out, L.id_in, L.id_out, (int)W, 0, 24, stream));
  a.order = L.id_out;
  return MP_OK;
}

// Every walker onto the stiff queue: a spec the explicit kernel does not implement (Bucciantini torque).
__global__ void queue_all_kernel(int* queue, int* count, int W, const int* ids = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) queue[i] = ids ? ids[i] : i;
  if (i == 0) *count = W;
}

template <int MODE>
static int launch_eval(mp_handle* h, KernelArgs& a, cudaStream_t stream, int lane = -1) {
  if (a.W == 0) return MP_OK;
  int rc = prepare_queue(h, a, a.W, stream, lane);
  if (rc) return rc;
  if (a.sp.bucciantini) {
    a.resume = nullptr;                                  // every walker starts in the implicit kernel
    queue_all_kernel<<<(a.W + 255) / 256, 256, 0, stream>>>(a.queue, a.queue_count, a.W);
    eval_stiff_kernel<MODE, 64><<<(a.W + 63) / 64, 64, 0, stream>>>(a);
    MP_CUDA(cudaGetLastError());
    return MP_OK;
  }
  if (h->bucketing && MODE != kModeCurves && a.W >= 2048 && (rc = bucket_walkers(h, a, stream, lane))) return rc;
  if (MODE == kModeCurves) {
    // Curve output is dominated by the luminosity stage at up to 10 001 nodes, which a warp works through
    // one walker at a time with one node per lane.  With few walkers (a full-grid launch is bounded by its
    // 240 KB of output per walker) 32 walkers per warp leave most schedulers empty, so the walkers are
    // spread one per 2..32 lanes until the launch holds ~16 warps per SM (measured: 16 384 full-grid curves,
    // 3.5 warps per SM and 9 % FP64 pipe with dense warps).
    int lpw = 1;
    while (lpw < 32 && (long long)a.W * lpw * 2 <= (long long)h->sm_count * 16 * 32) lpw *= 2;
    a.lanes_per_walker = lpw;
    const long long threads = (long long)a.W * lpw;
    eval_kernel<MODE, 32><<<(unsigned)((threads + 31) / 32), 32, 0, stream>>>(a);
    eval_stiff_kernel<MODE, 32><<<stiff_grid(h, a.W, 32), 32, 0, stream>>>(a);
    MP_CUDA(cudaGetLastError());
    return MP_OK;
  }
  // small ensembles: 32-thread blocks spread the warps over more SMs
  if (a.W <= h->sm_count * 64 * 4) {
    eval_kernel<MODE, 32><<<(a.W + 31) / 32, 32, 0, stream>>>(a);
    eval_stiff_kernel<MODE, 32><<<stiff_grid(h, a.W, 32), 32, 0, stream>>>(a);
  } else {
    eval_kernel<MODE, 64><<<(a.W + 63) / 64, 64, 0, stream>>>(a);
    eval_stiff_kernel<MODE, 64><<<stiff_grid(h, a.W, 64), 64, 0, stream>>>(a);
  }
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}

extern "C" int mp_lnprob_batch_device(mp_handle* h, const double* d_theta, int32_t W, int32_t ndim,
                                      double* d_lnp, int32_t* d_status, int32_t* d_n_rhs, void* stream) {
  if (!h || (W > 0 && (!d_theta || !d_lnp))) return fail(MP_ERR_BAD_ARG, "mp_lnprob_batch_device: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  KernelArgs a;
  int rc = fill_args(h, a, h->data_nodes, true, ndim, W, true);
  if (rc) return rc;
  a.theta = d_theta;
  a.lnp = d_lnp;
  a.status = d_status;
  a.n_rhs = d_n_rhs;
  return launch_eval<kModeLnprob>(h, a, (cudaStream_t)stream);
}

// One full wave of the explicit kernel: 8 resident 64-thread blocks per SM.
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
  // Large batches go through in wave-sized chunks on two alternating lanes: the H2D copy of chunk
  // k+1 and the D2H copy of chunk k-1 overlap the kernel of chunk k (when the caller's buffers are
  // pinned; pageable buffers still work, the copies just serialise).
  const int wave = wave_walkers(h);
  const int chunk = (W <= wave + wave / 2) ? W : wave;
  KernelArgs a;
  if ((rc = fill_args(h, a, h->data_nodes, true, ndim, W, true))) return rc;
  int k = 0;
  for (int c0 = 0; c0 < W; c0 += chunk, ++k) {
    const int lane = k & 1;
    cudaStream_t st = h->lanes[lane].stream;
    int n = W - c0;
    if (n > chunk + chunk / 2) n = chunk;          // the last chunk absorbs a remainder below half a wave
    MP_CUDA(cudaMemcpyAsync(h->s_theta + (size_t)c0 * ndim, theta + (size_t)c0 * ndim, (size_t)n * ndim * sizeof(double),
                            cudaMemcpyHostToDevice, st));
    a.W = n;
    a.theta = h->s_theta + (size_t)c0 * ndim;
    a.lnp = h->s_lnp + c0;
    a.status = h->s_status + c0;
    a.n_rhs = h->s_nrhs + c0;
    if ((rc = launch_eval<kModeLnprob>(h, a, st, lane))) return rc;
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
  KernelArgs a;
  int rc = fill_args(h, a, h->data_nodes, true, ndim, W, false);
  if (rc) return rc;
  a.sp.unlog_mask = 0;  // model_lum takes physical parameters
  if ((rc = ensure(&h->s_theta, &h->cap_theta, (size_t)W * MP_MAX_NDIM))) return rc;
  if ((rc = ensure(&h->s_out, &h->cap_out, (size_t)W * h->D))) return rc;
  size_t capw = h->cap_w;
  if (W > (int)capw) {
    cudaFree(h->s_status);
    h->s_status = nullptr;
    cudaFree(h->s_lnp); cudaFree(h->s_nrhs);
    h->s_lnp = nullptr; h->s_nrhs = nullptr; h->cap_w = 0;
    MP_CUDA(cudaMalloc((void**)&h->s_lnp, (size_t)W * sizeof(double)));
    MP_CUDA(cudaMalloc((void**)&h->s_status, (size_t)W * sizeof(int)));
    MP_CUDA(cudaMalloc((void**)&h->s_nrhs, (size_t)W * sizeof(int)));
    h->cap_w = W;
  }
  MP_CUDA(cudaMemcpyAsync(h->s_theta, pars, (size_t)W * ndim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  a.theta = h->s_theta;
  a.out = h->s_out;
  a.status = h->s_status;
  if ((rc = launch_eval<kModeModelAtData>(h, a, h->stream, 0))) return rc;
  MP_CUDA(cudaMemcpyAsync(out, h->s_out, (size_t)W * h->D * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (status) MP_CUDA(cudaMemcpyAsync(status, h->s_status, (size_t)W * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
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

extern "C" int mp_model_curves_device(mp_handle* h, const double* d_pars, int32_t W, int32_t ndim,
                                      int32_t node_stride, double* d_out, double* d_state,
                                      int32_t* d_status, void* stream) {
  if (!h || (W > 0 && (!d_pars || !d_out))) return fail(MP_ERR_BAD_ARG, "mp_model_curves_device: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  DeviceNodes* dn = nullptr;
  int rc = curve_node_set(h, node_stride, &dn);
  if (rc) return rc;
  KernelArgs a;
  if ((rc = fill_args(h, a, *dn, false, ndim, W, false))) return rc;
  a.sp.unlog_mask = 0;
  a.theta = d_pars;
  a.out = d_out;
  a.state = d_state;
  a.status = d_status;
  return launch_eval<kModeCurves>(h, a, (cudaStream_t)stream);
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
  rc = mp_model_curves_device(h, h->s_theta, W, ndim, node_stride, h->s_out, state ? h->s_state : nullptr,
                              d_status, h->stream);
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

static int launch_stretch(mp_handle* h, StretchArgs& s, cudaStream_t stream) {
  const int n_active = s.n_active;
  if (n_active == 0) return MP_OK;
  int rc;
  if ((rc = prepare_queue(h, s.k, n_active, stream))) return rc;
  if (s.k.sp.bucciantini) {
    s.k.resume = nullptr;
    queue_all_kernel<<<(n_active + 255) / 256, 256, 0, stream>>>(s.k.queue, s.k.queue_count, n_active);
    stretch_stiff_kernel<64><<<(n_active + 63) / 64, 64, 0, stream>>>(s);
    MP_CUDA(cudaGetLastError());
    return MP_OK;
  }
  if (n_active <= h->sm_count * 64 * 4) {
    stretch_kernel<32><<<(n_active + 31) / 32, 32, 0, stream>>>(s);
    stretch_stiff_kernel<32><<<stiff_grid(h, n_active, 32), 32, 0, stream>>>(s);
  } else {
    stretch_kernel<64><<<(n_active + 63) / 64, 64, 0, stream>>>(s);
    stretch_stiff_kernel<64><<<stiff_grid(h, n_active, 64), 64, 0, stream>>>(s);
  }
  MP_CUDA(cudaGetLastError());
  return MP_OK;
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
  StretchArgs s;
  std::memset(&s, 0, sizeof(s));
  int rc = fill_args(h, s.k, h->data_nodes, true, ndim, n_active, true);
  if (rc) return rc;
  s.k.n_rhs = d_n_rhs;
  s.coords = d_coords;
  s.lnp = d_lnp;
  s.active = d_active;
  s.complement = d_complement;
  s.n_active = n_active;
  s.n_complement = n_complement;
  s.a = a;
  s.seed = seed;
  s.step = step;
  s.accepted = d_accepted;
  return launch_stretch(h, s, (cudaStream_t)stream);
}

static int check_ensemble(const mp_ensemble* e, const char* who) {
  if (!e || !e->coords || !e->lnp) return fail(MP_ERR_BAD_ARG, std::string(who) + ": null ensemble / coords / lnp");
  if (e->nwalkers < 2 || (e->nwalkers & 1)) return fail(MP_ERR_BAD_ARG, std::string(who) + ": nwalkers must be even");
  if (e->world < 1 || e->rank < 0 || e->rank >= e->world || (e->nwalkers / 2) % e->world)
    return fail(MP_ERR_BAD_ARG, std::string(who) + ": half-ensemble does not split evenly over the ranks");
  if (e->n_peers < 0 || e->n_peers > MP_MAX_PEERS) return fail(MP_ERR_BAD_ARG, std::string(who) + ": bad n_peers");
  return MP_OK;
}

extern "C" int mp_ensemble_half_step(mp_handle* h, const mp_ensemble* e, uint64_t step, int32_t split, void* stream) {
  if (!h) return fail(MP_ERR_BAD_ARG, "mp_ensemble_half_step: null handle");
  int rc = check_ensemble(e, "mp_ensemble_half_step");
  if (rc) return rc;
  if (split != 0 && split != 1) return fail(MP_ERR_BAD_ARG, "mp_ensemble_half_step: split must be 0 or 1");
  MP_CUDA(cudaSetDevice(h->device));
  const int half = e->nwalkers / 2, m = half / e->world;
  StretchArgs s;
  std::memset(&s, 0, sizeof(s));
  if ((rc = fill_args(h, s.k, h->data_nodes, true, e->ndim, m, true))) return rc;
  s.k.n_rhs = e->n_rhs;
  s.coords = e->coords;
  s.lnp = e->lnp;
  s.perm = make_split_perm(e->nwalkers, e->seed, step, e->randomize_split);
  s.pos0 = split * half + e->rank * m;
  s.cpos0 = (1 - split) * half;
  s.n_active = m;
  s.n_complement = half;
  s.a = e->a;
  s.seed = e->seed;
  s.step = 2 * step + (uint64_t)split;
  s.accepted = e->accepted;
  s.status = e->status;
  s.n_peers = e->n_peers;
  for (int p = 0; p < e->n_peers; ++p) {
    if (!e->peer_coords[p] || !e->peer_lnp[p]) return fail(MP_ERR_BAD_ARG, "mp_ensemble_half_step: null peer replica");
    s.peer_coords[p] = e->peer_coords[p];
    s.peer_lnp[p] = e->peer_lnp[p];
  }
  s.pack_out = e->pack_out;
  if (e->bad_rows && e->bad_count && e->bad_capacity > 0) {
    s.bad_rows = e->bad_rows;
    s.bad_count = e->bad_count;
    s.bad_capacity = e->bad_capacity;
  }
  return launch_stretch(h, s, (cudaStream_t)stream);
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

extern "C" int mp_last_stiff_count(mp_handle* h, int32_t* count) {
  if (!h || !count) return fail(MP_ERR_BAD_ARG, "mp_last_stiff_count: null pointer");
  MP_CUDA(cudaSetDevice(h->device));
  MP_CUDA(cudaDeviceSynchronize());
  const mp_handle::Lane* L = h->last_lane ? h->last_lane : &h->lanes[0];
  MP_CUDA(cudaMemcpy(count, L->queue_count, sizeof(int), cudaMemcpyDeviceToHost));
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
