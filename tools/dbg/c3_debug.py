"""Debug: where does the C3 full-size scene diverge from the oracle?  (run on the GPU box)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import substrata_b200 as sb
import harness
from harness import BigScene, FULL_SCENES, rel_rms

def run(tag, **ov):
    kw = dict(FULL_SCENES["c3"]); kw.update(ov)
    sc = BigScene(**kw)
    eng = sb.MixEngine(sc.F, sh_hrir=sc.hrir, max_sources=sc.n)
    sel = np.arange(sc.n)
    ids = sc.build(eng, sel)
    gouts, gbus = [], []
    for b in range(sc.nb):
        sc.step(eng, ids, sel, b)
        v, o = eng.mix()
        gouts.append(o.copy()); gbus.append(eng.ambisonic_bus().astype(np.float64))
    P = 8
    res = [None] * P
    import threading
    def work(p):
        be = sc.oracle()
        s2 = np.arange(p, sc.n, P)
        i2 = sc.build(be, s2)
        outs, buses = [], []
        for b in range(sc.nb):
            sc.step(be, i2, s2, b)
            v, o = be.mix()
            outs.append(o.copy()); buses.append(be.ambisonic_bus().astype(np.float64))
        res[p] = (np.stack(outs), np.stack(buses))
    th = [threading.Thread(target=work, args=(p,)) for p in range(P)]
    [t.start() for t in th]; [t.join() for t in th]
    oo = np.sum([r[0].astype(np.float64) for r in res], axis=0)
    ob = np.sum([r[1] for r in res], axis=0)
    go = np.stack(gouts); gb = np.stack(gbus)
    print(tag, "out rel_rms per buffer", [float("%.3g" % rel_rms(go[b], oo[b])) for b in range(sc.nb)],
          "bus rel_rms per buffer", [float("%.3g" % rel_rms(gb[b], ob[b])) for b in range(sc.nb)], flush=True)
    # where in the buffer is the bus error?
    for b in range(sc.nb):
        e = np.abs(gb[b] - ob[b]).max(axis=0)
        worst = np.argsort(e)[-3:]
        print("   b%d max|bus err| %.3g at frames %s ; err energy by frame quarter %s" % (b, e.max(), worst.tolist(),
              [float("%.3g" % np.sum((gb[b] - ob[b])[:, q * sc.F // 4:(q + 1) * sc.F // 4] ** 2)) for q in range(4)]), flush=True)

for tag, ov in [("n2048", dict(n=2048)), ("n8192", dict(n=8192)), ("n65536_notoggle", dict(volume_toggle=False)), ("n65536", dict()),
                ("n65536_F512", dict(frames=512)), ("n16384_still", dict(n=16384, volume_toggle=False))]:
    run(tag, **ov)
