"""Registers / stack / static shared memory of the hot kernels of the shipped library -> profiles/r02_final_resource_usage.md
(`python tools/resource_usage.py`; runs cuobjdump --dump-resource-usage on arrayfire_b200/libafb200.so, no GPU needed)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "arrayfire_b200", "libafb200.so")
WANT = [("pipe_scan_kernel<0, float, float, false", "C1 scan f32 (af_accum)"),
        ("pipe_scan_kernel<0, int, int, true, unsigned int", "C4 scan_by_key s32, u32 keys"),
        ("pipe_where_kernel<unsigned char", "C5 where u8"), ("pipe_where_kernel<float", "where f32"),
        ("span_kernel<afb::ValPolicy<0, float, float>", "C3 / C2 sum f32"), ("span_kernel<afb::ArgPolicy<3, float", "C3 imax f32"),
        ("span_kernel<afb::ValPolicy<4, unsigned char", "C5 count u8"), ("dim_kernel<afb::ValPolicy<0, float, float>, 4", "C2 sum dim 1 f32"),
        ("dim_scan_kernel<0, float, float", "scan along dim > 0 f32"), ("rbk_tile_kernel<0, int, int", "sum_by_key s32"),
        ("rbk_count_kernel", "by-key run count"), ("expr_reduce_kernel<float>", "fused expression reduce f32 (pi)"),
        ("expr_reduce_kernel<double>", "fused expression reduce f64 / integer operands"),
        ("span_kernel<afb::VarPolicy<float, float, float>", "var f32"), ("span_kernel<afb::MeanPolicy<float", "mean f32"),
        ("tile_kernel<unsigned int>", "af_tile 4-byte")]


def main():
    text = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    rows = [(m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)))
            for m in re.finditer(r"Function (\S+):\s*\n?\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", text)]
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = ["# r02 final tree -- registers / stack / static shared memory of the hot kernels (`cuobjdump --dump-resource-usage "
           "arrayfire_b200/libafb200.so`, sm_100a, CUDA 12.9; `tools/resource_usage.py`)", "",
           "%d kernels in the library; %d of them use a stack frame (local arrays or spills). One representative instantiation per hot kernel:"
           % (len(rows), sum(1 for r in rows if r[2] > 0)), "",
           "| kernel (demangled prefix) | serves | registers | stack bytes | static shared bytes |", "|---|---|---|---|---|"]
    for pref, what in WANT:
        hit = [(n, r) for n, r in zip(names, rows) if pref in n]
        if hit:
            n, r = hit[0]
            out.append("| `%s` | %s | %d | %d | %d |" % (n.split("(")[0][:110], what, r[1], r[2], r[3]))
    out += ["", "Dynamic shared memory (the pipelined kernels' stages, the expression stack) is set at launch and not part of these figures.", ""]
    with open(os.path.join(ROOT, "profiles", "r02_final_resource_usage.md"), "w") as f:
        f.write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
