#!/usr/bin/env python
"""cProfile of one warm Assembly.update(pb, "all") of the REAL fedoo routed to the kernels (fedoo_b200.install)."""
import cProfile
import io
import os
import pstats
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fedoo as fd  # noqa: E402
import fedoo_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
fedoo_b200.install(fd)
fd.ModelingSpace("3D")
mesh = fd.mesh.box_mesh(nx=n + 1, ny=n + 1, nz=n + 1, elm_type="hex8", name="Domain")
fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
fd.weakform.StressEquilibrium("law", name="wf")
asm = fd.Assembly.create("wf", "Domain", "hex8", name="A")
pb = fd.problem.Linear("A")
pb.set_X(np.random.default_rng(0).standard_normal(pb.n_dof) * 1e-3)
asm.update(pb, compute="all")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
asm.update(pb, compute="all")
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
print(s.getvalue()[:7000])
