"""How long is a launch of the explicit integrator for a given ORDER of its work list?  (developer tool, CPU only)

The per-walker step counts come from the per-walker core compiled for the host (tests/hostsim, `hs_cost_trace`);
the scheduler of `advance_kernel` (persistent warps, per-lane refill from one queue, refill patience) is replayed on
them with every warp advancing one trip per tick.  Prints, per ordering, the makespan in trips, the mean lanes per
trip and a model time (3.0 us per trip for a warp alone on its scheduler, 3.9 us at full residency).

    python tools/sched_sim.py [W] [prior|SIGMA] [DATASET]        e.g.  python tools/sched_sim.py 262144 0.2 Classic

What it showed (2^18 walkers, sigma = 0.2 around the Classic truth; ideal = 271 trips): the (epsilon, M*delta)
key of round 2 ends after 430 trips -- the long integrations start whenever their epsilon bin comes up, the last of
them at the very end, and the launch ends with a tail as long as one of them -- M*delta DESCENDING as the major
key after 331, the true longest-first order after 289.  On the GPU: 2.56 -> 2.27 ms.  (The trip counts the tool
predicts for the key in use agree with the kernel's own: 911 866 warp-trips at 28.2 lanes against 920 000 at 27.9
counted on the device.)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def cost_trace(th, dataset):
    import __graft_entry__ as g
    g.build()
    from magprop_b200 import _capi as A
    from magprop_b200.engine import time_grid
    from magprop_b200.synthetic.mcmc_eqns import lower as LO, upper as HI
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "_hostsim.so"))
    gd = np.load(os.path.join(ROOT, "tests", "golden", "lnprob_script.npz"))
    x, y, e = (np.ascontiguousarray(gd[f"{dataset}_{k}"]) for k in ("x", "y", "yerr"))
    spec = A.script_model_spec(); pr = A.prior_spec(LO, HI); grid = time_grid(None)
    W = th.shape[0]
    cs = np.zeros((W, 1), np.int32); tot = np.zeros(W, np.int32); dfr = np.zeros(W, np.int32)
    ss = np.zeros(W, np.int32); sr = np.zeros(W, np.int32)
    rc = hs.hs_cost_trace(C.byref(spec), C.byref(pr), A.ptr(grid), grid.size, A.ptr(x), A.ptr(y), A.ptr(e), x.size,
                          A.ptr(th), W, 6, 1 << 20, 1, A.ptr(cs), A.ptr(tot), A.ptr(dfr), A.ptr(ss), A.ptr(sr))
    assert rc == 0, rc
    return tot, dfr, ss


def simulate(cost, n_warps=2960, patience=16):
    n = cost.size; nxt = 0
    rem = np.zeros((n_warps, 32), np.int32); waited = np.zeros(n_warps, np.int32); empty = np.zeros(n_warps, bool)
    t = 0; warp_trips = 0; time = 0.0
    while True:
        idle = rem == 0
        nidle = idle.sum(1)
        waited = np.where(nidle > 0, waited + 1, 0)
        want = (nidle > 0) & ~empty & ((nidle == 32) | (waited > patience))
        if want.any():
            ws = np.nonzero(want)[0]
            cnt = nidle[ws]; base = nxt + np.concatenate(([0], np.cumsum(cnt)[:-1]))
            nxt += int(cnt.sum())
            for w, b, c in zip(ws, base, cnt):
                if b + c >= n: empty[w] = True
                take = max(0, min(c, n - b))
                if take > 0:
                    rem[w, np.nonzero(idle[w])[0][:take]] = cost[b:b + take]
                waited[w] = 0
        live = (rem > 0).any(1)
        if not live.any() and nxt >= n: break
        na = int(live.sum())
        warp_trips += na
        time += 3.0 + 0.9 * na / n_warps
        rem = np.maximum(rem - 1, 0)
        t += 1
    return t, warp_trips, time


def key_round2(th):
    ie = np.clip((th[:, 4] + 4) * 4, 0, 31).astype(int); im = np.clip((th[:, 2] + th[:, 5] + 10) * 24, 0, 255).astype(int)
    return ie * 256 + np.where(ie & 1, 255 - im, im)


def key_in_use(th):      # order_key_of() in magprop_kernels.cu
    ie = np.clip((th[:, 4] + 4) * 4, 0, 31).astype(int); im = np.clip((th[:, 2] + th[:, 5] + 10) * 24, 0, 255).astype(int)
    return (255 - im) * 32 + np.where(im & 1, 31 - ie, ie)


if __name__ == "__main__":
    from magprop_b200.synthetic.mcmc_eqns import lower as LO, upper as HI
    from magprop_b200.synthetic.synth_mcmc import truths
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    mode = sys.argv[2] if len(sys.argv) > 2 else "0.2"
    ds = sys.argv[3] if len(sys.argv) > 3 else "Classic"
    rng = np.random.RandomState(5)
    th = rng.uniform(LO, HI, size=(W, 6)) if mode == "prior" else np.clip(truths[ds] + float(mode) * rng.randn(W, 6), LO, HI)
    th = np.ascontiguousarray(th)
    tot, dfr, ss = cost_trace(th, ds)
    tot = np.maximum(tot, 1)
    n_warps = max(1, min(2960, W // 32))
    md = th[:, 2] + th[:, 5]
    print(f"{W} walkers ({mode}, {ds}): explicit steps mean {tot.mean():.1f} max {tot.max()}, handed to the implicit integrator "
          f"{(dfr >= 0).mean():.3f}; corr(M*delta, log steps) {np.corrcoef(md, np.log(tot))[0, 1]:.3f}; ideal {tot.sum() / n_warps / 32:.0f} trips")
    orders = {"as given": np.arange(W), "round-2 key (epsilon major)": np.argsort(key_round2(th), kind="stable"),
              "key in use (M*delta descending major)": np.argsort(key_in_use(th), kind="stable"),
              "longest first (oracle)": np.argsort(-tot, kind="stable")}
    for nm, o in orders.items():
        t, wt, tm = simulate(tot[o], n_warps=n_warps)
        print(f"   {nm:40s} makespan {t:5d} trips  lanes/trip {tot.sum() / wt:5.2f}  model time {tm / 1e3:.3f} ms")
