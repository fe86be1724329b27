#!/bin/bash
# Alternate variants built by tools/ab_build.sh on the box this runs on:
#   tools/ab_run.sh ROUNDS NAME1 NAME2 [...] -- command...
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
rounds="$1"; shift
names=()
while [ "$1" != "--" ]; do names+=("$1"); shift; done
shift
cp "$root/vit-with-opencl_b200/libvit_b200.so" "$root/vit-with-opencl_b200/build/ab/.default.so"
for r in $(seq "$rounds"); do
  for n in "${names[@]}"; do
    cp "$root/vit-with-opencl_b200/build/ab/$n.so" "$root/vit-with-opencl_b200/libvit_b200.so"
    echo "== round $r, variant $n"
    "$@"
  done
done
cp "$root/vit-with-opencl_b200/build/ab/.default.so" "$root/vit-with-opencl_b200/libvit_b200.so"
