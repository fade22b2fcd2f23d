#!/bin/bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference sources where they lie under /root/reference (never copied
# into this repo) together with oracle/ref_harness.cpp into oracle/_ref/libref.so.
# Flags follow the reference's build.sh:10-49 (-std=c++20 -O3 -D_GNU_SOURCE -DDISABLE_NUMA,
# every header directory on -I); its Makefile is never used (it has no -O flag).
#
# Architecture flags: build.sh picks "-march=native -mavx512f -mavx512bw -mavx512vl -mavx512dq"
# on an AVX-512 host.  The AVX-512 kernels are dead code (macro typo __AVX_512F__,
# x86_simd.cpp:131 ...), so the live path is AVX2 + FMA contraction.  We therefore build
#   libref.so         with -march=haswell (AVX2+FMA; runs on any GPU-box host CPU)
#   libref_native.so  with build.sh's exact flags (only when the build host has avx512f)
# and tests/test_oracle_ref.py checks the two are bit-identical on every pinned vector.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${FASTLLAMA_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box) - keeping prebuilt files" >&2
    exit 0
fi
mkdir -p "$OUT"
SRCS=$(find "$REF/src" -name '*.cpp' ! -name 'main.cpp' | sort | tr '\n' ' ')
INCS=$(find "$REF/src" \( -name '*.h' -o -name '*.hpp' \) -exec dirname {} \; | sort -u | sed 's/^/-I/' | tr '\n' ' ')
COMMON="-std=c++20 -O3 -D_GNU_SOURCE -DDISABLE_NUMA -fPIC -shared -w"
build() { # $1 = output, $2.. = arch flags
    local out="$1"; shift
    if [ "$out" -nt "$HERE/ref_harness.cpp" ] && [ "$out" -nt "$HERE/build_ref.sh" ]; then
        echo "up to date: $out"; return
    fi
    echo "g++ -> $out ($*)"
    g++ -o "$out" $SRCS "$HERE/ref_harness.cpp" $COMMON "$@" $INCS -lpthread -lm
}
build "$OUT/libref.so" -march=haswell &
if grep -qi '^flags.*\<avx512f\>' /proc/cpuinfo; then
    build "$OUT/libref_native.so" -march=native -mavx512f -mavx512bw -mavx512vl -mavx512dq &
fi
# The reference-side binding of include/fastllama_b200.h, compiled against the reference's headers (tests/cxx/fl_binding.cpp):
# reference loader -> fl_upload / fl_forward.  Links nothing of the product: the library is dlopen'ed at run time.
BIND="$HERE/../tests/cxx/fl_binding.cpp"
if [ ! "$OUT/libfl_binding.so" -nt "$BIND" ] || [ ! "$OUT/libfl_binding.so" -nt "$HERE/../include/fastllama_b200.h" ]; then
    echo "g++ -> $OUT/libfl_binding.so"
    g++ -o "$OUT/libfl_binding.so" $SRCS "$BIND" $COMMON -march=haswell $INCS -lpthread -lm -ldl &
fi
wait
# The reference CLI itself, for the CPU baseline (bench.py --impl reference can also use libref.so).
if [ ! -x "$OUT/main" ] || [ "$OUT/main" -ot "$HERE/build_ref.sh" ]; then
    g++ -o "$OUT/main" $(find "$REF/src" -name '*.cpp' | sort | tr '\n' ' ') -std=c++20 -O3 -D_GNU_SOURCE -DDISABLE_NUMA -w -march=haswell $INCS -lpthread -lm
fi
ls -la "$OUT"
