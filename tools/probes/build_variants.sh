#!/bin/bash
# Builds A/B variants of libsparse_b200.so that differ only in -D flags of csrc/attention.cu (resident-CTA caps).
# usage: tools/probes/build_variants.sh name "-DSB200_ATT_FWD_CTAS32=5 ..." [name flags ...]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CS="$ROOT/opensearch-sparse-model-tuning-sample_b200/csrc"
OUT="$ROOT/tools/probes/variants"
mkdir -p "$OUT"
OBJS=$(ls "$CS"/build/*.o | grep -v attention.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       $flags -c "$CS/attention.cu" -o "$OUT/attention_$name.o"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libsparse_b200_$name.so" $OBJS "$OUT/attention_$name.o" -cudart static
  rm "$OUT/attention_$name.o"
  echo "built $OUT/libsparse_b200_$name.so"
done
