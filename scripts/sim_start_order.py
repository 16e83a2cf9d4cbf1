"""Why the starts of a multistart are handed out longest-first: replay of the slot scheduling of the
rollout kernel (W start slots in lock-step rounds, a finished slot is refilled from the start queue) on the per-start evaluation
counts the CPU oracle records, for several hand-out orders. CPU only.   python scripts/sim_start_order.py [C3] [trajectories] [W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
from oracle import oracle as orc


def rounds(e, order, W):
    """lock-step rounds until every start of `order` (evaluations e[i] each) has finished on W slots"""
    q = list(order); t = 0
    rem = [e[q.pop(0)] for _ in range(min(W, len(q)))]
    while rem:
        t += 1
        nxt = []
        for x in rem:
            if x > 1: nxt.append(x - 1)
            elif q: nxt.append(e[q.pop(0)])
        rem = nxt
    return t


def main(name="C3", M=48, W=5):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name, M=M)
    sur = wl.surrogate()
    rn = orc.gen_low_discrepancy_sequence(M, wl.d, wl.h + 1)
    starts = orc.generate_initial_guesses(wl.S, wl.lbs, wl.ubs)
    r = g._oracle_problem(orc, wl, sur, rn, starts, 0).rollout(tape=True)
    ev = (np.asarray(r["start_iters"]) + 1 + (np.asarray(r["start_status"]) != 6)).transpose(2, 1, 0)  # [M][h][S] evaluations per start
    S = ev.shape[2]
    tot = {k: 0 for k in ("natural order (round-2 build before)", "previous step of the same trajectory", "kernel rule: previous step; step 1 from the previous trajectory's step 1",
                          "true longest-first (not available in advance)", "lower bound max(sum / W, longest start)")}
    first = None
    for m in range(M):
        prev = None
        for j in range(ev.shape[1]):
            e = ev[m, j]
            lpt = lambda key: list(np.argsort(-key, kind="stable")) if key is not None else list(range(S))
            tot["natural order (round-2 build before)"] += rounds(e, range(S), W)
            tot["previous step of the same trajectory"] += rounds(e, lpt(prev), W)
            tot["kernel rule: previous step; step 1 from the previous trajectory's step 1"] += rounds(e, lpt(first if j == 0 else prev), W)
            tot["true longest-first (not available in advance)"] += rounds(e, lpt(e), W)
            tot["lower bound max(sum / W, longest start)"] += max(int(np.ceil(e.sum() / W)), int(e.max()))
            if j == 0: first = e
            prev = e
    print(f"{name}: {M} trajectories of the CPU oracle, {S} starts, W = {W} slots, {ev.mean():.2f} evaluations per start; lock-step rounds per trajectory:")
    for k, v in tot.items():
        print(f"  {k:<78} {v / M:7.1f}")
    print("  mean evaluations per (step, start):")
    print(np.round(ev.mean(axis=0), 1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C3", int(sys.argv[2]) if len(sys.argv) > 2 else 48, int(sys.argv[3]) if len(sys.argv) > 3 else 5)
