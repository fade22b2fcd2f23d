"""CPU: the index arithmetic the kernels share, compiled for the host with nvcc (no GPU needed) and checked exhaustively:
  * k_cache_index (csrc/kernels.cuh): the K cache row layout is a permutation of the row, lane j's q-th float4 sits at float4
    index 8 q + j and holds the reference's AVX lane j in chain order (elements 8 i + j, i = 4 q .. 4 q + 3; x86_simd.cpp:1447-1468);
  * lane_elem (csrc/megakernel.cuh): the 8 lanes of a quantisation group own disjoint pairs that cover the group, and in every
    load the 8 lanes read 8 consecutive pairs (whole 32-byte sectors of tagged words)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

SRC = r'''
#include <cstdio>
#include <vector>
#include "megakernel.cuh"
using namespace fl;
template <int GS> static int check_lane_elem() {
    constexpr int PER = GS / 8, LPP = PER / 2;
    std::vector<int> seen(GS, 0);
    for (int sub = 0; sub < 8; ++sub)
        for (int q = 0; q < LPP; ++q) {
            const int e = lane_elem<GS>(sub, q);
            if (e < 0 || e + 1 >= GS + 0 + 1 || (e & 1)) return 1;
            seen[e]++; seen[e + 1]++;
        }
    for (int e = 0; e < GS; ++e) if (seen[e] != 1) return 2;
    if (kPollIL)
        for (int q = 0; q < LPP; ++q)
            for (int sub = 0; sub < 8; ++sub) if (lane_elem<GS>(sub, q) != 16 * q + 2 * sub) return 3;      // 8 lanes x 16 bytes contiguous
    return 0;
}
int main() {
    for (int hs : {64, 128}) {
        std::vector<int> seen(hs, 0);
        for (int e = 0; e < hs; ++e) { const int s = k_cache_index(e); if (s < 0 || s >= hs) return 10; seen[s]++; }
        for (int s = 0; s < hs; ++s) if (seen[s] != 1) return 11;
        for (int j = 0; j < 8; ++j)
            for (int i = 0; i < hs / 8; ++i)
                if (k_cache_index(8 * i + j) != ((i / 4) * 8 + j) * 4 + i % 4) return 12;
    }
    if (int rc = check_lane_elem<64>()) return 20 + rc;
    if (int rc = check_lane_elem<32>()) return 30 + rc;
    std::printf("ok\n");
    return 0;
}
'''


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_shared_index_arithmetic_on_the_host(tmp_path):
    src = tmp_path / "layout_host.cu"
    src.write_text(SRC)
    exe = tmp_path / "layout_host"
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-fmad=false", "-I", os.path.join(ROOT, "fast-llama_b200", "csrc"),
           "-o", str(exe), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)
