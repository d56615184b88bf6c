#!/bin/bash
# Builds an A/B variant of libmvg_b200.so with extra -D flags:
#   tools/build_variant.sh <name> "<extra nvcc flags>"  -> mvgformer_b200/variants/libmvg_<name>.so
# Select it at run time with MVG_LIB_PATH (mvgformer_b200/_lib.py).
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/mvgformer_b200/csrc
out=$root/mvgformer_b200/variants
bld=$src/build_$name
mkdir -p $out $bld
flags="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for f in api deform_forward deform_backward pyramid project_sample select_pad offsets_dlt elementwise linear_tcgen05 pre_post ffn_chain cameras decoder_driver offset_chain; do
  if [ "$f" = "project_sample" ] || [ "$f" = "${VARIANT_SRC:-project_sample}" ] || [ ! -f $src/build/$f.o ]; then
    nvcc $flags $extra -c ${PS_SRC:-$src/$f.cu} -o $bld/$f.o &
  else
    cp $src/build/$f.o $bld/$f.o
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmvg_$name.so $bld/*.o -lcudart_static -ldl
rm -rf $bld
echo built $out/libmvg_$name.so
