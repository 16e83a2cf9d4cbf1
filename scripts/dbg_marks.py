"""Debug aid: launches the rollout asynchronously with a marker-instrumented build and prints where each warp is after a few seconds."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g

def main(name, M):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name, M=M)
    sur = wl.surrogate()
    eng = pkg.RolloutEngine(0)
    eng.set_surrogate(pkg.FantasySurrogate(sur, wl.h))
    eng.generate_normals(M, wl.h + 1)
    eng.set_starts(pkg.generate_initial_guesses(wl.S, wl.lbs, wl.ubs))
    eng.rollout_device(wl.x0, wl.theta, wl.lbs, wl.ubs, wl.h, float(np.min(sur.y)), 0)
    time.sleep(4.0)
    out = (ctypes.c_int * (160 * 32))()
    rc = eng.lib.rbo_debug_marks(out)
    print("marks rc", rc)
    a = np.array(out[:]).reshape(160, 32)
    import collections
    cnt = collections.Counter()
    for b in range(min(M, 148)):
        cnt[(tuple(a[b, :16]), tuple(a[b, 16:]))] += 1
    for (reach, passed), n in cnt.most_common(6):
        print(n, "blocks: reached", reach, "\n          passed ", passed)
    sys.stdout.flush()
    os._exit(0)

main(sys.argv[1], int(sys.argv[2]))
