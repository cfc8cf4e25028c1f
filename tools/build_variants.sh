#!/bin/bash
# builds A/B variants of libvecb200.so HERE (no GPU time): tools/build_variants.sh name "DEFS" [name "DEFS" ...]
# each lands in pgvector_b200/variants/libvecb200_<name>.so (git-ignored, travels with gpurun); the default build is restored last
set -e
cd "$(dirname "$0")/.."
mkdir -p pgvector_b200/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
  VB_NVCC_DEFS="$defs" python -m pgvector_b200.build > /dev/null
  cp pgvector_b200/libvecb200.so pgvector_b200/variants/libvecb200_$name.so
  echo "built $name ($defs)"
done
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
python -m pgvector_b200.build > /dev/null
cp pgvector_b200/libvecb200.so pgvector_b200/variants/libvecb200_default.so
echo "default restored"
