#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the reference's own volume-rendering CUDA ops, UNMODIFIED,
# from the sources where they lie under /root/reference, into oracle/_ref/libvolrend_ref.so.
#
# The reference path (deps/volume-rendering-jax/lib/impl/{marching,integrating,packbits}.cu)
# compiles from its own three source files; we do not run its CMake/Nix build.  Nothing is copied
# into this repo: only the compiled .so lands in oracle/_ref/ (git-ignored, travels to the GPU box).
# jax-tcnn (the reference's inference hash encoder) needs tiny-cuda-nn v1.6, which is not vendored
# (deps/_sources/generated.nix:45-55) -> unbuildable here, see DESIGN.md.
#
# The .so exports C++-mangled volrendjax::{march_rays,...} with the XLA legacy custom-call
# signature (cudaStream_t, void**, const char*, size_t); tests/refops.py resolves them.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NGP_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/deps/volume-rendering-jax/lib/impl"
if [ ! -d "$SRC" ]; then
    echo "build_ref: $SRC not present (GPU box?) -- keeping prebuilt $OUT" >&2
    exit 0
fi
TORCH_INC="$(python - <<'EOF'
import os, torch
print(os.path.join(os.path.dirname(torch.__file__), "include"))
EOF
)"
mkdir -p "$OUT"
FLAGS=(-std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
       -w -DFMT_HEADER_ONLY -include stdexcept -include string -include cstring -include cstdint
       -I"$REF/deps" -I"$TORCH_INC" -Xcompiler -fPIC)
for f in marching integrating packbits; do
    nvcc "${FLAGS[@]}" -c "$SRC/$f.cu" -o "$OUT/$f.o"
done
nvcc -shared -o "$OUT/libvolrend_ref.so" "$OUT/marching.o" "$OUT/integrating.o" "$OUT/packbits.o"
rm -f "$OUT"/*.o
echo "build_ref: wrote $OUT/libvolrend_ref.so"
