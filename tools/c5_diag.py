import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms, meshgen
from tools.bench_configs import rhs3, timed
torch.set_default_dtype(torch.float64)
meshes, data = meshgen.seven_fracture_network(1024, 586)
with torch.device("cuda"):
    mesh = tfem.FracturesTri(meshes, torch.tensor(data))
    basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 3))
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
load = forms.Load(rhs3)
t0=time.time(); plan = basis.tile_plan(elem_ids=True); torch.cuda.synchronize(); print("plan s", time.time()-t0, "tiles", plan.n_tiles, "templates", plan.n_templates, "index MB", plan.index_bytes/1e6, "max_elem", plan.max_elem, "halo", plan.halo_factor, "lattice", plan.lattice)
for path in ("tiled", "two_pass"):
    t = timed(lambda: basis.assemble(forms.Stiffness(), load, layout="values", path=path), 10, flush)
    print(path, round(t,1), "us")
