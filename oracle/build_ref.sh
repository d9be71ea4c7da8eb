#!/usr/bin/env bash
# ORACLE tooling: compile the reference's own hash-grid CUDA operator, UNMODIFIED, from the sources where they
# lie under /root/reference, into oracle/_ref/_gridencoder_ref.so (git-ignored; travels to the GPU box).
# It is the GPU-side check of the hash-grid restatement (tests/test_hashgrid_gpu.py) and the
# "beat THAT kernel" baseline in bench.py --impl refkernel.  Only difference from the reference's own build
# (gridencoder/backend.py:6-9): -std=c++17 (torch 2.11 headers reject c++14) and an explicit sm_100a target.
set -euo pipefail
REF=${OCCNERF_REFERENCE_ROOT:-/root/reference}/core/nets/occnerf/gridencoder/src
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
[ -f "$REF/gridencoder.cu" ] || { echo "reference sources not present; keeping prebuilt $OUT"; exit 0; }
PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'PYEOF'
import torch.utils.cpp_extension as c, sysconfig
print(" ".join("-I"+p for p in c.include_paths() + [sysconfig.get_paths()["include"]]))
PYEOF
)
TORCH_LIB=$($PY -c 'import torch,os;print(os.path.join(os.path.dirname(torch.__file__),"lib"))')
ABI=$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')
nvcc -O3 -std=c++17 -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ \
  -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=$ABI \
  -DTORCH_EXTENSION_NAME=_gridencoder_ref -DTORCH_API_INCLUDE_EXTENSION_H $TORCH_INC \
  "$REF/gridencoder.cu" "$REF/bindings.cpp" \
  -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python -Xlinker -rpath -Xlinker "$TORCH_LIB" \
  -o "$OUT/_gridencoder_ref.so"
echo "built $OUT/_gridencoder_ref.so"
