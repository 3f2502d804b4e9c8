"""Build tests/emu/_build/lib<name>_emu.so: one pna_b200/csrc/*.cu with every kernel launch rewritten into a sequential host
loop (cuda_host_shim.h), compiled by g++.  Test infrastructure for the CPU suite: kernels without intra-block communication
(the backward, the halo pull) can be run thread after thread, which checks their index arithmetic and control flow.  Inline
PTX (only in kernels that are not emulated, e.g. the flag barrier) is replaced by a call that aborts.

Memory checking (a host-side stand-in for compute-sanitizer): set PNA_EMU_ASAN=1 and preload the sanitizer runtime,
    PNA_EMU_ASAN=1 ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
        python -m pytest tests/test_bwd_emulated.py tests/test_pull_emulated.py -q
then every access of the emulated kernels into the torch-allocated buffers is bounds-checked (round 2: clean)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pna_b200", "csrc")
BUILD = os.path.join(HERE, "_build")


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


def strip_inline_ptx(text: str) -> str:
    """`asm [volatile](...);` statements -> emu_unsupported(): the kernels that contain them are not emulated."""
    out, pos = "", 0
    for m in re.finditer(r"\basm\s*(volatile\s*)?\(", text):
        if m.start() < pos:
            continue
        depth, k = 0, m.end() - 1
        while True:
            depth += text[k] == "("
            depth -= text[k] == ")"
            if depth == 0:
                break
            k += 1
        out += text[pos:m.start()] + "emu_unsupported()"
        pos = k + 1
    return out + text[pos:]


def build(source: str = "pna_aggregate_bwd.cu", force: bool = False, asan: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    asan = asan or os.environ.get("PNA_EMU_ASAN") == "1"
    src = os.path.join(CSRC, source)
    stem = os.path.splitext(source)[0]
    lib = os.path.join(BUILD, f"lib{stem}_emu{'_asan' if asan else ''}.so")
    deps = [src, os.path.join(HERE, "cuda_host_shim.h"), __file__, os.path.join(CSRC, "pna_aggregate.cuh"),
            os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "pna_b200.h")]
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    body = strip_inline_ptx(rewrite_launches(open(src).read()))
    body = re.sub(r'#include "(pna_aggregate\.cuh|common\.cuh)"', lambda m: f'#include "{CSRC}/{m.group(1)}"', body)
    tu = os.path.join(BUILD, f"{stem}_emu.cpp")
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
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", f"-I{cuda_inc}", tu, "-o", lib]
    if asan:      # out-of-bounds accesses of the kernels into the (torch-allocated) buffers: see the module docstring
        cmd[1:1] = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["pna_aggregate_bwd.cu"]
    for name in names:
        print(build(name, force="--force" in sys.argv))
