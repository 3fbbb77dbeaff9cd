#!/bin/bash
# developer A/B run on the GPU box: tools/ab_run.sh <out.jsonl> <lib1.so> <lib2.so> ...   (env GGP_BS, GGP_CHAINS, GGP_STEPS pass through)
out=$1; shift
mkdir -p "$(dirname "$out")"
: > "$out"
for lib in "$@"; do
  GGP_LIB=$(realpath "$lib") timeout 300 python tools/ab_bench.py 2> "gpurun_out/ab_$(basename "$lib").err" | tail -1 >> "$out" || echo "{\"lib\": \"$lib\", \"failed\": true}" >> "$out"
done
cat "$out"
