"""the headline workload (hex8 box n^3, ElasticIsotrop, random U) reported as K-only, R-only and K+R, on the
regular and on the jittered mesh (interior nodes displaced by U(-0.2h, 0.2h): no two elements alike).  Device-resident
timing with CUDA events, 3 warm-up + 8 timed launches each.  One JSON line per mesh.

    python scripts/variants_bench.py --n 200
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fedoo_b200 as fd  # noqa: E402
from fedoo_b200 import dist as fdist  # noqa: E402


def timed(fn, warm=3, steps=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    a = ap.parse_args()
    n = a.n
    for jitter in (False, True):
        fd.Assembly.delete_memory()
        loc = fdist.box_local_slab(n, 0, 1, jitter=jitter)
        fd.ModelingSpace("3D")
        fd.Mesh(loc.nodes, loc.elements, "hex8", name="Domain")
        fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="ElasticLaw")
        fd.weakform.StressEquilibrium("ElasticLaw", name="weakform")
        asm = fd.Assembly.create("weakform", "Domain", "hex8", name="Assembling", reuse_buffers=True, vector_on_device=True)
        pb = fd.problem.Linear("Assembling")
        U = torch.from_numpy(np.random.default_rng(0).standard_normal(3 * (n + 1) ** 3) * 1e-3).cuda()
        pb.set_X(U)
        asm.update(pb, compute="all")
        out = {"workload": f"hex8 box {n}^3, ElasticIsotrop, jitter={jitter}", "n_elems": n**3}
        for label, compute in (("K+R", "all"), ("K", "matrix"), ("R", "vector")):
            ms = timed(lambda: asm.assemble_global_mat(compute))
            out[label] = {"ms": round(ms, 3), "Melem/s": round(n**3 / ms / 1e3, 1)}
        print(json.dumps(out))
        del asm, pb, U
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
