// hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles magprop_core.cuh for the host
// (g++, no CUDA) so the per-walker algorithm can be debugged against the oracle
// in the GPU-less build container.  Never linked into or loaded by the package;
// the product path is the CUDA library only.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../magprop_b200/csrc/magprop_host.hpp"
#include "../../magprop_b200/csrc/magprop_rng.cuh"

using namespace mp;

static DataView view_of(const NodeProgram& np, double t_start) {
  DataView dv;
  dv.n_nodes = (int)np.node_t.size();
  dv.n_data = (int)np.y.size();
  dv.node_t = np.node_t.data();
  dv.dat_ys = np.ys.data();
  dv.dat_c = np.c.data();
  dv.dat_dx = np.dx.data();
  dv.dat_w = np.w.data();
  dv.dat_lo = np.lo.data();
  dv.t_start = t_start;
  return dv;
}

// The kernels' three stages for one walker on the host: setup, advance (explicit integrator, handing over to
// the implicit one when the walker turns stiff), reduce.  `theta` is what the caller of the corresponding
// entry point passes (log-space for lnprob, physical for the model calls -- sp.unlog_mask says which).
template <int MODE>
static double evaluate_staged(const Spec& sp, const DataView& dv, const double* theta, int ndim, const mp_prior_spec* pr,
                              int& st, int& nr, double* out, double* state, const int* dat_orig) {
  const int Nn = dv.n_nodes;
  const double t_end = dv.node_t[Nn - 1];
  WalkerRec r;
  prepare_walker(sp, theta, ndim, pr && pr->enabled, pr ? pr->lower : nullptr, pr ? pr->upper : nullptr, dv.t_start, t_end, r);
  st = r.status;
  nr = r.n_rhs;
  if (st & kWalkerPriorReject) return 0.0;
  std::vector<double> row(Nn, NAN);
  if (st == kWalkerOk) {
    Integrator in;
    int jn = 0;
    bool stiff = false;
    StiffRec q;
    if (!sp.bucciantini) {
      integrator_load(r, r.w.C, dv.t_start, in);
      jn = drain_nodes<false>(in, 0, Nn, dv.node_t, row.data());
      while (jn < Nn && in.status == kWalkerOk && !in.stiff) {
        integrator_step(sp, r.w, t_end, in);
        jn = drain_nodes<false>(in, jn, Nn, dv.node_t, row.data());
      }
      if (jn < Nn && in.status == kWalkerOk && in.stiff) {
        q.t = in.t; q.y = in.omega; q.h = in.h; q.wid = 0; q.jn = jn; q.n_rhs = in.n_rhs; q.n_steps = in.n_steps;
        stiff = true;
      }
    } else {
      q.t = dv.t_start; q.y = r.y0; q.h = r.h0; q.wid = 0; q.jn = 0; q.n_rhs = r.n_rhs; q.n_steps = 0;
      stiff = true;
    }
    if (stiff) {
      integrator_load_stiff(q, in);
      jn = drain_nodes<true>(in, q.jn, Nn, dv.node_t, row.data());
      while (jn < Nn && in.status == kWalkerOk) {
        radau_step(sp, r.w, t_end, in);
        jn = drain_nodes<true>(in, jn, Nn, dv.node_t, row.data());
      }
    }
    st |= in.status;
    nr = in.n_rhs;
    for (; jn < Nn; ++jn) row[jn] = NAN;
  }
  return reduce_rows<MODE>(sp, dv, r.w, row.data(), out, state, 1, dat_orig);
}

extern "C" int hs_lnprob(const mp_model_spec* ms, const mp_prior_spec* pr, const double* grid, int G,
                         const double* t, const double* y, const double* yerr, int D,
                         const double* theta, int W, int ndim, double* lnp, int* status, int* nrhs,
                         int* nsteps) {
  NodeProgram np;
  int rc = build_node_program(grid, G, t, y, yerr, D, np);
  if (rc) return rc;
  Spec sp = make_spec(*ms);
  DataView dv = view_of(np, grid[0]);
  for (int w = 0; w < W; ++w) {
    int st = 0, nr = 0;
    const double chi2 = evaluate_staged<kModeLnprob>(sp, dv, theta + (size_t)w * ndim, ndim, pr, st, nr, nullptr, nullptr, nullptr);
    lnp[w] = (st & kWalkerPriorReject) ? -INFINITY : lnlike_of(chi2, st);
    if (status) status[w] = st;
    if (nrhs) nrhs[w] = nr;
  }
  (void)nsteps;
  return 0;
}

extern "C" int hs_curves(const mp_model_spec* ms, const double* grid, int G, const double* pars_in,
                         int W, int ndim, int stride, double* out, double* state, int* status,
                         int* nrhs) {
  std::vector<double> node_t;
  std::vector<int> gi;
  build_curve_nodes(grid, G, stride, node_t, gi);
  NodeProgram np;
  np.node_t = node_t;
  Spec sp = make_spec(*ms);
  sp.unlog_mask = 0;
  DataView dv = view_of(np, grid[0]);
  const int Gs = (int)node_t.size();
  for (int w = 0; w < W; ++w) {
    int st = 0, nr = 0;
    evaluate_staged<kModeCurves>(sp, dv, pars_in + (size_t)w * ndim, ndim, nullptr, st, nr, out + (size_t)w * 3 * Gs,
                                 state ? state + (size_t)w * 2 * Gs : nullptr, nullptr);
    if (status) status[w] = st;
    if (nrhs) nrhs[w] = nr;
  }
  return Gs;
}

extern "C" double hs_disc_S(double u) { return disc_S(u); }

// node program of a dataset, for tests of the host-side preparation
extern "C" int hs_node_program(const double* grid, int G, const double* t, const double* y, const double* yerr,
                               int D, int* n_nodes, int* node_grid_index, int* lo, double* dx, double* Dx,
                               int* order) {
  NodeProgram np;
  int rc = build_node_program(grid, G, t, y, yerr, D, np);
  if (rc) return rc;
  *n_nodes = (int)np.node_t.size();
  for (size_t i = 0; i < np.node_t.size(); ++i) node_grid_index[i] = np.node_grid_index[i];
  for (int i = 0; i < D; ++i) { lo[i] = np.lo[i]; dx[i] = np.dx[i]; Dx[i] = np.Dx[i]; order[i] = np.order[i]; }
  return 0;
}

extern "C" int hs_model_at(const mp_model_spec* ms, const double* grid, int G, const double* t, int D,
                           const double* pars_in, int W, int ndim, double* out, int* status) {
  std::vector<double> ones(D, 1.0);
  NodeProgram np;
  int rc = build_node_program(grid, G, t, ones.data(), ones.data(), D, np);
  if (rc) return rc;
  Spec sp = make_spec(*ms);
  sp.unlog_mask = 0;
  DataView dv = view_of(np, grid[0]);
  for (int w = 0; w < W; ++w) {
    int st = 0, nr = 0;
    evaluate_staged<kModeModelAtData>(sp, dv, pars_in + (size_t)w * ndim, ndim, nullptr, st, nr, out + (size_t)w * D, nullptr,
                                      np.order.data());
    if (status) status[w] = st;
  }
  return 0;
}

// per-step trace of one walker (debug aid): rows of (t, h, omega after the step, accepted, implicit)
extern "C" int hs_trace(const mp_model_spec* ms, const double* grid, int G, const double* pars_in, double t_end,
                        int implicit, double* out, int max_rows) {
  (void)G;
  Spec sp = make_spec(*ms);
  sp.unlog_mask = 0;
  double pars[6], de, pe, fb;
  unpack_theta(sp, pars_in, 6, pars, de, pe, fb);
  Walker wk;
  walker_setup(sp, pars, de, pe, fb, grid[0], wk);
  Integrator in;
  in.n_rhs = 0;
  if (implicit) integrator_init<false>(sp, wk, grid[0], t_end, in);
  else integrator_init<true>(sp, wk, grid[0], t_end, in);   // explicit variant: state is omega^-2
  int n = 0;
  while (in.t < t_end && in.status == 0 && n < max_rows && !(in.stiff && !implicit)) {
    const double t0 = in.t, h0 = in.h;
    if (implicit) radau_step(sp, wk, t_end, in);
    else integrator_step(sp, wk, t_end, in);
    double* r = out + (size_t)n * 5;
    r[0] = t0; r[1] = h0; r[2] = implicit ? in.omega : 1.0 / std::sqrt(in.omega); r[3] = in.t > t0; r[4] = implicit;
    ++n;
  }
  return n;
}

// the ensemble order (keyed permutation) as the kernels compute it
extern "C" void hs_ensemble_order(int n, unsigned long long seed, unsigned long long step, int randomize, int* order) {
  const SplitPerm p = make_split_perm(n, seed, step, randomize);
  for (int g = 0; g < n; ++g) order[g] = (int)perm_at(p, (uint32_t)g);
}
extern "C" int hs_ensemble_order_roundtrip(int n, unsigned long long seed, unsigned long long step) {
  const SplitPerm p = make_split_perm(n, seed, step, 1);
  for (int g = 0; g < n; ++g)
    if ((int)perm_inv(p, perm_at(p, (uint32_t)g)) != g) return g + 1;
  return 0;
}

// Cost trace of the explicit pass (design aid for the launch scheduling): for each walker, the number of
// step attempts it makes inside each chunk of NB nodes, whether / after how many attempts it is deferred as
// stiff, and the implicit pass's step and RHS counts from the hand-over point.
extern "C" int hs_cost_trace(const mp_model_spec* ms, const mp_prior_spec* pr, const double* grid, int G,
                             const double* t, const double* y, const double* yerr, int D, const double* theta,
                             int W, int ndim, int NB, int max_chunks, int* chunk_steps /*[W][max_chunks]*/,
                             int* total_steps, int* deferred_at, int* stiff_steps, int* stiff_rhs) {
  NodeProgram np;
  int rc = build_node_program(grid, G, t, y, yerr, D, np);
  if (rc) return rc;
  Spec sp = make_spec(*ms);
  const int Nn = (int)np.node_t.size();
  const double t_end = np.node_t[Nn - 1];
  for (int w = 0; w < W; ++w) {
    const double* th = theta + (size_t)w * ndim;
    int* cs = chunk_steps + (size_t)w * max_chunks;
    for (int c = 0; c < max_chunks; ++c) cs[c] = 0;
    total_steps[w] = 0; deferred_at[w] = -1; stiff_steps[w] = 0; stiff_rhs[w] = 0;
    if (pr->enabled && !prior_accepts(th, ndim, pr->lower, pr->upper)) continue;
    double pars[6], de, pe, fb;
    unpack_theta(sp, th, ndim, pars, de, pe, fb);
    Walker wk;
    walker_setup(sp, pars, de, pe, fb, grid[0], wk);
    if (wk.bad) continue;
    Integrator in;
    in.n_rhs = 0;
    integrator_init<true>(sp, wk, grid[0], t_end, in);
    int jn = 0;
    bool deferred = false;
    for (int c0 = 0, c = 0; c0 < Nn && !deferred; c0 += NB, ++c) {
      const int c1 = (c0 + NB < Nn) ? c0 + NB : Nn;
      while (jn < c1) {
        while (jn < c1 && np.node_t[jn] <= in.t) ++jn;
        if (jn >= c1) break;
        if (in.status != kWalkerOk) { jn = c1; break; }
        if (in.stiff) { deferred = true; break; }
        integrator_step(sp, wk, t_end, in);
        if (c < max_chunks) cs[c]++;
        total_steps[w]++;
      }
      if (in.status != kWalkerOk) break;
    }
    if (deferred) {
      deferred_at[w] = total_steps[w];
      Integrator im;
      integrator_resume(in.t, 1.0 / std::sqrt(in.omega), in.h, im);
      while (im.t < t_end && im.status == kWalkerOk) {
        radau_step(sp, wk, t_end, im);
        stiff_steps[w]++;
      }
      stiff_rhs[w] = im.n_rhs;
    }
  }
  return 0;
}

// the tables' local coordinate and row index of u (fast = the explicit integrator's table), for the test of the
// mantissa-shift form against the centre / scale definition
extern "C" int hs_table_coord(double u, int fast, double* s, int* row, int* nsub_log2) {
  TableAt ta;
  const bool in = fast ? table_locate_fast(u, ta) : table_locate_safe(u, ta);
  *s = ta.s;
  *row = fast ? (int)((ta.row - &mp_disc_fast[0][0]) / (long)(sizeof(mp_disc_fast[0]) / sizeof(double)))
              : (int)((ta.row - &mp_disc_table[0][0]) / (long)(sizeof(mp_disc_table[0]) / sizeof(double)));
  *nsub_log2 = fast ? MP_DISC_FAST_NSUB_LOG2 : MP_DISC_NSUB_LOG2;
  return in ? 1 : 0;
}

// the step-size controller: 1/fac after an accepted (accept != 0) or a rejected step, and the updated log2(facold)
extern "C" double hs_step_scale(double aerr, double sk, float lfacold, int accept, float* lfacold_new) {
  const float lerr = log2_error_ratio(aerr, sk);
  if (accept) {
    const float g = StepControl::expo1 * lerr - StepControl::beta * lfacold - StepControl::l_safe;
    *lfacold_new = fmaxf(lerr, StepControl::l_facold_min);
    return step_scale(-fmaxf(StepControl::l_grow, fminf(StepControl::l_shrink, g)));
  }
  *lfacold_new = lfacold;
  return step_scale(-fminf(StepControl::l_shrink, StepControl::expo1 * lerr - StepControl::l_safe));
}
