"""Build tests/emu/_build/libpna_bwd_emu.so: pna_b200/csrc/pna_aggregate_bwd.cu with every kernel launch rewritten into a
sequential host loop (cuda_host_shim.h), compiled by g++.  Test infrastructure for the CPU suite: the backward kernels have
no intra-block communication, so running their threads one after the other checks index arithmetic and control flow."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "pna_b200", "csrc", "pna_aggregate_bwd.cu")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libpna_bwd_emu.so")


def _split_top_level(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> str:
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            return out + text[pos:]
        line_start = text.rfind("\n", 0, i) + 1
        kern = text[line_start:i].strip()
        j = text.index(">>>", i)
        cfg = _split_top_level(text[i + 3:j])
        assert text[j + 3] == "(", text[j:j + 40]
        depth, k = 0, j + 3
        while True:
            depth += text[k] == "("
            depth -= text[k] == ")"
            if depth == 0:
                break
            k += 1
        args = text[j + 4:k]
        indent = text[line_start:i][: len(text[line_start:i]) - len(text[line_start:i].lstrip())]
        out += text[pos:line_start] + f"{indent}EMU_LAUNCH(({kern}), ({cfg[0]}), ({cfg[1]}), {args})"
        pos = k + 1


def build(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    deps = [SRC, os.path.join(HERE, "cuda_host_shim.h"), __file__,
            os.path.join(ROOT, "pna_b200", "csrc", "pna_aggregate.cuh"), os.path.join(ROOT, "pna_b200", "csrc", "common.cuh"),
            os.path.join(ROOT, "include", "pna_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    body = rewrite_launches(open(SRC).read())
    body = body.replace('#include "pna_aggregate.cuh"', f'#include "{ROOT}/pna_b200/csrc/pna_aggregate.cuh"')
    tu = os.path.join(BUILD, "pna_aggregate_bwd_emu.cpp")
    with open(tu, "w") as f:
        f.write(f'#include "{HERE}/cuda_host_shim.h"\n')
        f.write(body)
        f.write("""
#include <stdarg.h>
namespace pna {
static thread_local char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t, const char* what) { set_error("%s", what); return PNA_ERR_CUDA; }
}
extern "C" const char* emu_last_error(void) { return pna::g_err; }
""")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", f"-I{cuda_inc}", tu, "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
