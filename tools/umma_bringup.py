"""Bring-up probe for the tcgen05 pair kernel: compares it with the FP64 SIMT kernel on small
problems and prints where they differ.  Run under `timeout` on the GPU box."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import weightedld_b200 as wld
from weightedld_b200.synth import make_alignment


def run(chars, kernel, thr=-1.0, limbs=3, weights=None, ctas=2):
    with wld.Context(0) as ctx:
        ctx.set_pair_kernel(kernel)
        ctx.set_cta_group(ctas)
        ctx.set_limbs(limbs)
        ctx.load_alignment(chars)
        k = ctx.filter_sites()
        if weights is None:
            ctx.henikoff()
        else:
            ctx.set_weights(weights)
        n, done = ctx.ld_pairs(thr)
        return k, ctx.fetch_pairs(n, 1), done, ctx.stage_ms(wld.STAGE_PAIR), ctx.pair_info()


ok = True
for (n, l, limbs) in [(64, 100, 1), (64, 100, 3), (200, 300, 3), (1000, 900, 3), (1000, 900, 2), (1000, 900, 4), (3000, 3000, 3)]:
    chars = make_alignment(n, l, seed=n + l, block=60, clonal=True)
    w = np.ones(n, np.float32) if limbs == 1 else None
    k, s, sd, sms, _ = run(chars, "simt", limbs=limbs, weights=w)
    same = True
    for kern, ctas in (("bf16", 1), ("i8", 1), ("bf16", 2), ("i8", 2)):
        try:
            k2, u, ud, ums, info = run(chars, kern, limbs=limbs, weights=w, ctas=ctas)
        except Exception as e:  # noqa
            print(f"[{n}x{l} limbs={limbs}] {kern} x{ctas}cta FAILED: {e}")
            ok = False
            same = False
            break
        same = len(s) == len(u) and s.tobytes() == u.tobytes()
        print(f"[{n}x{l} limbs={limbs}] kept={k} simt: {len(s)} pairs {sms:.3f} ms | {kern} x{ctas}cta: {len(u)} pairs {ums:.3f} ms "
              f"tiles={info.tiles} limb_bits={info.limb_bits} done {sd}/{ud} -> {'IDENTICAL' if same else 'DIFFERENT'}")
        if not same:
            break
    if not same:
        ok = False
        sm = {(int(p['site_a']), int(p['site_b'])): p for p in s}
        um = {(int(p['site_a']), int(p['site_b'])): p for p in u}
        missing = sorted(set(sm) - set(um))[:10]
        extra = sorted(set(um) - set(sm))[:10]
        print("  missing in umma:", missing)
        print("  extra in umma:", extra)
        bad = [(k_, sm[k_], um[k_]) for k_ in sorted(set(sm) & set(um)) if sm[k_].tobytes() != um[k_].tobytes()][:10]
        for k_, a, b in bad:
            print("  differs", k_, a, b)
        break
print("BRINGUP", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
