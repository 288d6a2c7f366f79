"""GPU diagnosis tool: renders single triangles of the `pathological` scene family through the CUDA path (binned and direct)
and through the CPU oracle, and prints where they differ (vertex data, bounds of the differing pixels, first values).
Used to find the tile-candidate bug of polygons with out-of-range coordinates; edit CASES for other triangles.
usage (GPU box): python scripts/diag_pathological.py"""
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, dataclasses
import scenes
from oracle import swref
from swiftshader_b200.scene import Device, Frame, Scene
dev = Device(0)
def cuda(scene, binned):
    dev.set_option("force_binned", binned)
    fr = Frame(dev, scene)
    try:
        fr.upload_inputs(); fr.upload_attachments(); fr.draw(); fr.resolve(); fr.download_all(); dev.sync()
        att = {k: v.copy() for k, v in fr.att.items()}
        res = fr.resolved[0].copy() if fr.resolved is not None else None
    finally:
        fr.close()
    return scenes.outputs(scene, att, res)
seen = 0
CASES = {1: (21, 23), 0: (6,), 19: (35, 43), 3: (11, 35), 17: (37,)}
for seed in CASES:
    sc = scenes.pathological(seed)
    d = sc.draws[0]
    V = d.vertices.reshape(-1, 3, 8)
    for i in CASES[seed]:
        d1 = dataclasses.replace(d, vertices=np.ascontiguousarray(V[i]))
        s1 = dataclasses.replace(sc, draws=[d1])
        want = swref.render_oracle(s1)
        wres = swref.resolve_oracle(s1, want) if s1.samples > 1 else None
        wout = scenes.outputs(s1, want, wres)
        for binned in (1, 0):
            got = cuda(s1, binned)
            nz = {k: int((got[k].view(np.uint8) != wout[k].view(np.uint8)).sum()) for k in wout}
            if any(nz.values()):
                cov_w = int((wout["color"].reshape(-1, 4) != wout["color"].reshape(-1, 4)[0]).any(axis=1).sum())
                cov_g = int((got["color"].reshape(-1, 4) != wout["color"].reshape(-1, 4)[0]).any(axis=1).sum())
                print(f"seed {seed} tri {i} binned={binned} ms={s1.samples} diff {nz} oracle_cov {cov_w} cuda_cov {cov_g}")
                for k in wout:
                    g, w = got[k].reshape(sc.height if k != "x" else -1, sc.width, -1), wout[k].reshape(sc.height, sc.width, -1)
                    bad = np.argwhere((g.view(np.uint8) != w.view(np.uint8)).any(axis=2))
                    ys, xs = bad[:, 0], bad[:, 1]
                    if len(bad):
                        print(f"   {k}: rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()} first", [(int(y), int(x), g[y, x].tolist(), w[y, x].tolist()) for y, x in bad[:4]])
                seen += 1
    if seen > 24: break
print("done", seen)
