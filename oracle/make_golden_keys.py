"""ORACLE tooling: write tests/golden/state_dict_keys.json = name -> [shape, dtype] of the UNMODIFIED reference
Network's state_dict (core/nets/occnerf/network.py:29-88 + the per-subject parameters of generate_neural_points,
network.py:90-146), built on CPU through oracle/ref_shim.py.  This is what a reference checkpoint
(trainer.py:398-415, `{'network': state_dict}`) contains, so it pins the drop-in claim "reference checkpoints load".

    python -m oracle.make_golden_keys
"""
from __future__ import annotations

import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occnerf_b200 import synthetic as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "state_dict_keys.json")


SIG_PATH = os.path.join(os.path.dirname(PATH), "gridencoder_signatures.json")


def operator_signatures():
    """Parameter names, in order, of the three functions the reference's pybind module exports
    (gridencoder/src/gridencoder.h:12-15, bound by name in src/bindings.cpp:5-9), parsed from the header itself."""
    import re
    hdr = open(os.path.join(ref_shim.REF_ROOT, "core", "nets", "occnerf", "gridencoder", "src", "gridencoder.h")).read()
    bind = open(os.path.join(ref_shim.REF_ROOT, "core", "nets", "occnerf", "gridencoder", "src", "bindings.cpp")).read()
    out = {}
    for name, args in re.findall(r"void\s+(\w+)\s*\(([^;]*)\)\s*;", hdr):
        assert f'"{name}"' in bind, f"{name} is declared but not bound"
        out[name] = [a.strip().split()[-1] for a in args.split(",")]
    return out


API_PATH = os.path.join(os.path.dirname(PATH), "network_signatures.json")
API_METHODS = ("forward", "_render_rays", "_query_mlp", "_batchify_rays", "deploy_mlps_to_secondary_gpus")


def network_api():
    """Positional parameter names (and the **kwargs name) of the Network methods that stay as the API surface
    (network.py:542-549, 435-447, 164-170, 307, 149), read from the reference source with ast."""
    import ast
    src = open(os.path.join(ref_shim.REF_ROOT, "core", "nets", "occnerf", "network.py")).read()
    out = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == "Network":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in API_METHODS:
                    out[f.name] = {"args": [a.arg for a in f.args.args], "kwargs": f.args.kwarg.arg if f.args.kwarg else None}
    assert sorted(out) == sorted(API_METHODS)
    return out


def main():
    warnings.filterwarnings("ignore", category=FutureWarning)
    with open(API_PATH, "w") as f:
        json.dump(network_api(), f, indent=1, sort_keys=True)
    print(API_PATH)
    sigs = operator_signatures()
    with open(SIG_PATH, "w") as f:
        json.dump(sigs, f, indent=1, sort_keys=True)
    print(SIG_PATH, {k: len(v) for k, v in sigs.items()})
    sub = S.make_subject(seed=0)
    net = ref_shim.build_reference_network(sub, S.make_weights(sub.bound, seed=0))
    keys = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in net.state_dict().items()}
    enc = net.cnl_mlp.module.encoder                      # GridEncoder.__init__, grid.py:97-141: the level table
    grid = {"offsets": [int(o) for o in enc.offsets.tolist()], "per_level_scale": float(enc.per_level_scale),
            "base_resolution": int(enc.base_resolution), "num_levels": int(enc.num_levels), "level_dim": int(enc.level_dim),
            "input_dim": int(enc.input_dim), "n_params": int(enc.n_params), "bound": float(sub.bound)}
    with open(os.path.join(os.path.dirname(PATH), "hashgrid_levels.json"), "w") as f:
        json.dump(grid, f, indent=1, sort_keys=True)
    with open(PATH, "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)
    print(PATH, len(keys), "entries,", sum(int(__import__("math").prod(s)) for s, _ in keys.values()), "elements")


if __name__ == "__main__":
    main()
