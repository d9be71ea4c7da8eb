"""Candidate grid of occnerf_knn_grid on the synthetic subject: cells, list entries, bytes and build time per cell size."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occnerf_b200 import ops, synthetic
d = torch.device("cuda")
sub = synthetic.make_subject()
base = sub.point_base.to(d).float()
fps = [f.to(d) for f in sub.fps_index]
rows = []
for cell in [float(c) for c in (sys.argv[1:] or ["0.025", "0.0125"])]:
    torch.cuda.synchronize(); t0 = time.time()
    g = ops.build_knn_grid(base, fps, cell=cell)
    torch.cuda.synchronize()
    cnt = g["cell_tab"][:, :, 1].float()
    rows.append(dict(cell=cell, cells=g["cells"], entries=g["entries"], list_MB=round(g["entries"] * 2 / 2**20, 1),
                     cell_tab_MB=round(g["cells"] * 32 / 2**20, 1), build_s=round(time.time() - t0, 2),
                     mean_candidates_per_level=[round(float(x), 1) for x in cnt.mean(0)],
                     max_candidates_per_level=[int(x) for x in cnt.max(0)[0]]))
    ops._GRID_CACHE.clear(); del g
print(json.dumps(rows, indent=1))
