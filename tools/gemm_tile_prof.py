#!/usr/bin/env python
"""Per-tile timeline of the forward / backward-data tensor-core GEMM (csrc/gemm_tc.cuh gemm_tc_kernel) on train-step
shapes: where do the ~13-15 us per 128-row tile go?  (ncu of the final round-2 build: the K = 128 and the K = 256 launches
of the 256-wide layer take 94.6 / 106 us for the same 1024 tiles -- the main loop is not what bounds them.)

The product sources are NOT touched: `--build` (here, no GPU) copies csrc/ to scratch/prof_csrc/, inserts %globaltimer
stamps for CTA 0 into the copy (the string edits below are the whole instrumentation) and links scratch/libgcbf_prof.so;
`--run` (on the GPU box) loads that library through ctypes, runs the shapes and prints, per tile of CTA 0:

  tma0   first operand load of the tile issued           (TMA warp)
  tma1   last operand load of the tile issued
  mma0   MMA warp: accumulator free, first k-block's operand split done -> first MMA issued
  mma1   MMA warp: last MMA of the tile issued + committed
  epi0   epilogue warp: accumulator complete (tmem_full observed)
  epi1   epilogue warp: tile stored, accumulator released

usage:  python tools/gemm_tile_prof.py --build
        gpurun -- 'python tools/gemm_tile_prof.py --run > gpurun_out/r02_gemm_tile_prof.txt'
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gcbfplus_b200", "csrc")
DST = os.path.join(ROOT, "scratch", "prof_csrc")
LIB = os.path.join(ROOT, "scratch", "libgcbf_prof.so")
NT = 12   # tiles of CTA 0 that are stamped

EDITS = [
    # the buffer + stamp helper
    ("template <int BN, int EPI, bool ACCUM, int BKc>\n__global__ void __launch_bounds__(THREADS_NN, 1)\ngemm_tc_kernel(",
     "__device__ unsigned long long g_gemm_prof[%d * 8 + 8];\n"
     "__device__ __forceinline__ void prof_stamp(int tile, int slot) {\n"
     "    if (blockIdx.x == 0 && tile < %d) {\n"
     "        unsigned long long t;\n"
     "        asm volatile(\"mov.u64 %%0, %%%%globaltimer;\" : \"=l\"(t));\n"
     "        g_gemm_prof[tile * 8 + slot] = t;\n"
     "    }\n"
     "}\n"
     "template <int BN, int EPI, bool ACCUM, int BKc>\n__global__ void __launch_bounds__(THREADS_NN, 1)\ngemm_tc_kernel(" % (NT, NT)),
    # TMA warp
    ("            uint32_t it = 0;\n            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n"
     "                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;\n"
     "                for (int kb = 0; kb < nkb; ++kb, ++it) {\n"
     "                    const int s = it % STAGES;\n"
     "                    const uint32_t ph = (it / STAGES) & 1;\n"
     "                    mbar_wait(&empty[s], ph ^ 1);\n",
     "            uint32_t it = 0;\n            int ptile = 0;\n"
     "            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ptile) {\n"
     "                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;\n"
     "                for (int kb = 0; kb < nkb; ++kb, ++it) {\n"
     "                    const int s = it % STAGES;\n"
     "                    const uint32_t ph = (it / STAGES) & 1;\n"
     "                    mbar_wait(&empty[s], ph ^ 1);\n"
     "                    if (kb == 0) prof_stamp(ptile, 0);\n"
     "                    if (kb == nkb - 1) prof_stamp(ptile, 1);\n"),
    # MMA warp: first MMA of the tile / commit
    ("                    mbar_wait(&conv[s], ph);\n"
     "                    asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n"
     "                    const uint32_t a_hi = smem_u32(smem + s * CF::STAGE_BYTES);",
     "                    mbar_wait(&conv[s], ph);\n"
     "                    asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n"
     "                    if (kb == 0) prof_stamp((int)tcount, 2);\n"
     "                    const uint32_t a_hi = smem_u32(smem + s * CF::STAGE_BYTES);"),
    ("                umma_commit(&tmem_full[acc]);\n",
     "                umma_commit(&tmem_full[acc]);\n                prof_stamp((int)tcount, 3);\n"),
    # epilogue warp (first epilogue warp, lane 0)
    ("            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);\n"
     "            asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n"
     "            const int row = quarter * 32 + lane;",
     "            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);\n"
     "            asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n"
     "            if (warp == 6 && lane == 0) prof_stamp((int)tcount, 4);\n"
     "            const int row = quarter * 32 + lane;"),
    ("            asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n"
     "            mbar_arrive(&tmem_empty[acc]);\n",
     "            asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n"
     "            if (warp == 6 && lane == 0) prof_stamp((int)tcount, 5);\n"
     "            mbar_arrive(&tmem_empty[acc]);\n"),
]
GETTER = '''
extern "C" __attribute__((visibility("default"))) int32_t gcbf_gemm_prof_read(unsigned long long* host_out) {
    return (int32_t)cudaMemcpyFromSymbol(host_out, gcbf::tc::g_gemm_prof, sizeof(unsigned long long) * (%d * 8 + 8));
}
''' % NT


def build():
    sys.path.insert(0, ROOT)
    from gcbfplus_b200 import build as B
    if os.path.exists(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST)
    p = os.path.join(DST, "gemm_tc.cuh")
    s = open(p).read()
    for old, new in EDITS:
        assert s.count(old) == 1, old[:80]
        s = s.replace(old, new)
    open(p, "w").write(s)
    with open(os.path.join(DST, "gnn.cu"), "a") as f:
        f.write(GETTER)
    # the include path "../../include/gcbf_b200.h" of common.cuh is relative to csrc/: scratch/prof_csrc/../../include
    objs, procs = [], []
    for unit, extra in B.UNITS.items():
        obj = os.path.join(DST, unit.replace(".cu", ".o"))
        cmd = [B._nvcc(), *B.ARCH, *B.COMMON, *extra, "-c", os.path.join(DST, unit), "-o", obj]
        procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for unit, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise SystemExit(f"nvcc failed on {unit}:\n{out.decode()}")
    r = subprocess.run([B._nvcc(), *B.ARCH, "-shared", "-o", LIB, *objs, "-lcudart"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise SystemExit(r.stdout.decode())
    print("built", LIB)


def run():
    import torch
    lib = C.CDLL(LIB)
    P = C.c_void_p
    lib.gcbf_gemm_tc.restype = C.c_int32
    lib.gcbf_gemm_tc.argtypes = [C.c_int32, C.c_int32] + [P] * 8 + [C.c_int32] * 4 + [P]
    lib.gcbf_split_tf32.restype = C.c_int32
    lib.gcbf_split_tf32.argtypes = [P, P, P, C.c_int32, P]
    lib.gcbf_gemm_prof_read.restype = C.c_int32
    lib.gcbf_gemm_prof_read.argtypes = [P]
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    names = ["tma0", "tma1", "mma0", "mma1", "epi0", "epi1"]
    for (M, K, N, epi, what) in [(131072, 128, 256, 1, "update layer 1 (bias + ReLU), agents"),
                                 (131072, 256, 256, 1, "folded update/head layer (bias + ReLU), agents"),
                                 (185000, 128, 256, 3, "backward-data msg -> x1 (ReLU mask), edges"),
                                 (185000, 256, 128, 0, "folded message layer (bias), edges")]:
        A = torch.randn(M, K, device="cuda", generator=g)
        Bt = (torch.randn(N, K, device="cuda", generator=g) * 0.1).contiguous()
        Bh, Bl = torch.empty_like(Bt), torch.empty_like(Bt)
        assert lib.gcbf_split_tf32(Bt.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), Bt.numel(), st) == 0
        b = torch.randn(N, device="cuda", generator=g)
        aux = torch.randn(M, N, device="cuda", generator=g)
        out = torch.empty(M, N, device="cuda")

        def call():
            rc = lib.gcbf_gemm_tc(epi, 0, A.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), b.data_ptr(), None, out.data_ptr(),
                                  aux.data_ptr(), None, M, M, K, N, st)
            assert rc == 0, rc
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        buf = (C.c_ulonglong * (NT * 8 + 8))()
        assert lib.gcbf_gemm_prof_read(buf) == 0
        tiles = (M + 127) // 128 * (1 if N == 128 or 2 * ((M + 127) // 128) > 148 else 2)
        print(f"\n## M={M} K={K} N={N} epi={epi}: {what}\n{us:.1f} us per launch, {tiles} tiles over 148 CTAs "
              f"= {us / ((tiles + 147) // 148):.2f} us per tile round; {K // 32} k-blocks per tile")
        t00 = buf[0]
        print("tile | " + " | ".join(f"{n:>7s}" for n in names) + " |  (us since the first load of tile 0)")
        for t in range(min(NT, (tiles + 147) // 148)):
            row = [(buf[t * 8 + s] - t00) / 1e3 for s in range(6)]
            print(f"{t:4d} | " + " | ".join(f"{v:7.2f}" for v in row) + " |")


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    elif "--run" in sys.argv:
        run()
    else:
        print(__doc__)
