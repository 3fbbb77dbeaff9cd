#!/bin/bash
# developer: build several library variants in parallel.  tools/build_variants.sh name1:"flags" name2:"flags" ...  -> build_var/lib_<name>.so
cd "$(dirname "$0")/.."
mkdir -p build_var
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  ( tools/build_variant.sh build_var/lib_$n.so $f > build_var/build_$n.log 2>&1 && echo "built $n" || echo "FAILED $n" ) &
done
wait
