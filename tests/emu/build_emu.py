#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: compile the UNMODIFIED sources of sts_b200/csrc with g++ against tests/emu/cuda_emu.h into
tests/emu/_build/libstsgpu_emu.so, a functional CPU emulation of the engine for the `-m "not gpu"` suite.

Three textual rewrites are all g++ needs (the files under sts_b200/ are never touched):
  kernel<<<grid, block, smem, stream>>>(args)   ->  EMU_LAUNCH(kernel, grid, block, smem, stream, args)
  extern __shared__ T name[];                    ->  T *name = (T *) emu::dyn_smem();
  the three cp.async inline-PTX helpers of k_universal.cu -> a plain 4-byte copy / nothing
"""
import glob
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sts_b200", "csrc")
# STS_EMU_ASAN=1: a second build with AddressSanitizer (tests/emu/_build_asan), for hunting out-of-bounds accesses of
# device / pinned buffers by kernels and host code; run python with LD_PRELOAD=$(g++ -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
ASAN = os.environ.get("STS_EMU_ASAN", "") not in ("", "0")
# STS_EMU_UBSAN=1: the same with UndefinedBehaviorSanitizer (tests/emu/_build_ubsan): shift counts out of range,
# signed overflow, misaligned or out-of-bounds accesses - x86 and the GPU resolve these differently, so each report
# is a place where the emulator and the hardware may disagree
UBSAN = os.environ.get("STS_EMU_UBSAN", "") not in ("", "0")
BUILD = os.path.join(HERE, "_build_asan" if ASAN else "_build_ubsan" if UBSAN else "_build")
LIB = os.path.join(BUILD, "libstsgpu_emu.so")
CXX = os.environ.get("CXX", "g++")
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-w", "-Wno-unknown-pragmas",
         "-include", os.path.join(HERE, "cuda_emu.h"), "-I", os.path.join(HERE, "stub"), "-I", CSRC,
         "-I", os.path.join(ROOT, "include")] + (["-fsanitize=address", "-fno-omit-frame-pointer"] if ASAN else []) + \
    (["-fsanitize=undefined", "-fno-sanitize=vptr", "-fno-omit-frame-pointer"] if UBSAN else [])


def _balanced(s, i, open_ch, close_ch):
    """s[i] == open_ch: index just past the matching close_ch"""
    depth = 0
    while True:
        c = s[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite_launches(src):
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier, optionally followed by a balanced <...> template argument list
        j = k
        while src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while src[j - 1].isalnum() or src[j - 1] in "_:":
            j -= 1
        kernel = src[j:k].strip()
        e = src.index(">>>", k)
        cfg = src[k + 3:e]
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", "launch without argument list near %r" % src[k - 40:k + 40]
        b = _balanced(src, a, "(", ")")
        args = src[a + 1:b - 1]
        parts, depth, cur = [], 0, ""
        for c in cfg:                                   # split the launch configuration at top-level commas
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += c
        parts.append(cur.strip())
        while len(parts) < 4:
            parts.append("0")
        out.append(src[pos:j])
        out.append("EMU_LAUNCH((%s), %s, %s, %s, %s%s)" % (kernel, parts[0], parts[1], parts[2], parts[3],
                                                          (", " + args) if args.strip() else ""))
        pos = b


def rewrite(src):
    src = rewrite_launches(src)
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];",
                 r"\1 *\2 = (\1 *) emu::dyn_smem();", src)
    src = re.sub(r'asm volatile\("cp\.async\.ca\.shared\.global[^\n]*\n', "*smem_dst = *gsrc; (void) d;\n", src)
    src = re.sub(r'asm volatile\("cp\.async\.(?:commit_group|wait_group)[^\n]*?"memory"\);', ";", src)
    assert "<<<" not in src and "asm volatile" not in src
    return src


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "*.cpp")) + [os.path.abspath(__file__),
                                                                                          os.path.join(ROOT, "include", "stsgpu.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    jobs = []
    for s in srcs:
        name = os.path.splitext(os.path.basename(s))[0]
        cpp = os.path.join(BUILD, name + ".emu.cpp")
        with open(cpp, "w") as f:
            f.write('#line 1 "%s"\n' % s)
            f.write(rewrite(open(s).read()))
        extra = ["-O2"] if name.startswith("k_") else []
        jobs.append((cpp, os.path.join(BUILD, name + ".o"), extra))
    jobs.append((os.path.join(HERE, "cuda_emu_rt.cpp"), os.path.join(BUILD, "cuda_emu_rt.o"), ["-O2"]))

    def cc(job):
        src, obj, extra = job
        cmd = [CXX] + FLAGS + extra + ["-x", "c++", "-c", src, "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, r.returncode, r.stdout

    failed = False
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for src, rc, log in ex.map(cc, jobs):
            if rc != 0 or verbose:
                sys.stderr.write("== %s (rc %d)\n%s\n" % (src, rc, log[-6000:]))
            failed |= rc != 0
    if failed:
        raise SystemExit("cuda_emu build failed")
    subprocess.check_call([CXX, "-shared", "-pthread"] + (["-fsanitize=address"] if ASAN else []) + (["-fsanitize=undefined"] if UBSAN else []) + ["-o", LIB] + [j[1] for j in jobs] + ["-ldl", "-lm", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
