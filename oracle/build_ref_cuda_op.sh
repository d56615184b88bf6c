#!/bin/bash
# Builds the REFERENCE's own CUDA extension (lib/models/ops/src: vision.cpp, cpu/deform_cpu.cpp,
# cuda/deform_cuda.cu + deform_im2col_cuda.cuh) for sm_100a -> oracle/_ref/Deformable_ref*.so
# (git-ignored; it travels to the GPU box with the snapshot).  Test infrastructure: the kernel to
# beat in bench.py and a second oracle for mvg_deform_forward (tests/test_reference_cuda_op.py).
#
# The sources are compiled where they lie under $MVG_REFERENCE_ROOT (default /root/reference);
# nothing is copied into the repository.  The reference does not compile against torch >= 2.x as
# is: `AT_DISPATCH_FLOATING_TYPES(value.type(), ...)` (deform_cuda.cu:75,145) needs
# `value.scalar_type()`, and its setup.py stops at sm_75 / refuses to run without a GPU
# (lib/models/ops/setup.py:59-66).  The two-token fix is applied by sed into a scratch directory
# under $TMPDIR; every other line is the reference's.
set -euo pipefail
REF=${MVG_REFERENCE_ROOT:-/root/reference}
SRC=$REF/lib/models/ops/src
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$SRC" ] || { echo "reference sources not found at $SRC"; exit 3; }
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')
PYINC=$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')
EXT=$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')
ABI=$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/cuda" "$OUT"
sed 's/AT_DISPATCH_FLOATING_TYPES(value\.type()/AT_DISPATCH_FLOATING_TYPES(value.scalar_type()/' \
    "$SRC/cuda/deform_cuda.cu" > "$TMP/cuda/deform_cuda.cu"
grep -c 'value.scalar_type()' "$TMP/cuda/deform_cuda.cu" | grep -qx 2 || { echo "patch did not apply twice"; exit 4; }
INC="-I$SRC -I$SRC/cuda -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC"
DEF="-DWITH_CUDA -DTORCH_EXTENSION_NAME=Deformable_ref -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI"
NVCC=${NVCC:-nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC \
      -w $INC $DEF -c "$TMP/cuda/deform_cuda.cu" -o "$TMP/deform_cuda.o"
g++ -O2 -std=c++17 -fPIC -w $INC -I/usr/local/cuda/include $DEF -c "$SRC/vision.cpp" -o "$TMP/vision.o"
g++ -O2 -std=c++17 -fPIC -w $INC -I/usr/local/cuda/include $DEF -c "$SRC/cpu/deform_cpu.cpp" -o "$TMP/deform_cpu.o"
g++ -shared -o "$OUT/Deformable_ref$EXT" "$TMP/vision.o" "$TMP/deform_cpu.o" "$TMP/deform_cuda.o" \
    -L"$TORCH_DIR/lib" -Wl,-rpath,"$TORCH_DIR/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python \
    -L/usr/local/cuda/lib64 -lcudart
echo "built $OUT/Deformable_ref$EXT"
