import cProfile, pstats, time, sys, os
sys.path.insert(0, os.getcwd())
import torch
import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms, meshgen
mesh = meshgen.structured_rectangle(2048, 1024, jitter=0.25, seed=1234, topology=False)
with torch.device("cuda"):
    basis = tfem.Basis(tfem.MeshTri(mesh), tfem.ElementTri(1, 3))
pat = basis.pattern
coords_host = torch.from_numpy(mesh["vertices"]).pin_memory()
values_host = torch.empty(pat.nnz, dtype=torch.float64).pin_memory()
load_host = torch.empty(pat.n_dof, dtype=torch.float64).pin_memory()
bil, ld = forms.StiffnessMass(), forms.Load()
for i in range(6):
    t0 = time.perf_counter()
    basis.assemble_from_host(coords_host, bil, ld, values_host, load_host)
    torch.cuda.synchronize()
    print("call", i, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
pr = cProfile.Profile(); pr.enable()
basis.assemble_from_host(coords_host, bil, ld, values_host, load_host); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
