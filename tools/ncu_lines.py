"""Stall samples of a kernel in an ncu report, summed per SOURCE LINE (run here, no GPU needed):

    python tools/ncu_lines.py gpurun_out/r02_v9_step.ncu-rep rk45_attempt_kernelIdLb1ELi32ELb0 [top]

The ncu CSV source page lists SASS addresses only; the line table comes from `nvdisasm -g` of the cubin inside the
in-tree libfwb200.so (built with -lineinfo), so the library must be the build the capture was taken from.  Prints, per
line, the share of the kernel's stall samples and of its executed warp instructions."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tum_adlr_deep_reinforcement_learning_b200", "libfwb200.so")
CSRC = os.path.join(ROOT, "tum_adlr_deep_reinforcement_learning_b200", "csrc")


def line_table(mangled_fragment):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = os.path.join(tmp, "fw_step.sm_100a.cubin")
    text = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    table, cur, on = {}, None, False
    for ln in text.split("\n"):
        if ln.startswith("//-") and ".text." in ln:
            on = mangled_fragment in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, frag = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    table = line_table(frag)
    kname = re.sub(r"^_?[A-Za-z0-9]*?(rk45_[a-z]+_kernel|head_kernel).*", r"\1", frag)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = [r for r in rows if "Address" in r][0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
    seen, uniq = set(), []
    for r in data:                       # a report with several launches of the kernel repeats the listing
        if r[0] in seen:
            break
        seen.add(r[0])
        uniq.append(r)
    base = int(uniq[0][0], 16)
    smp, ins, tot_s, tot_i = collections.Counter(), collections.Counter(), 0, 0
    for r in uniq:
        k = table.get(int(r[0], 16) - base, ("?", 0))
        s, x = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        smp[k] += s; ins[k] += x; tot_s += s; tot_i += x
    src = {}
    print("%s: %d stall samples, %d warp instructions, %d SASS instructions mapped to lines" % (kname, tot_s, tot_i, len(table)))
    for (f, l), n in smp.most_common(top):
        if f not in src and os.path.exists(os.path.join(CSRC, f)):
            src[f] = open(os.path.join(CSRC, f)).read().split("\n")
        text = src[f][l - 1].strip()[:90] if f in src and 0 < l <= len(src[f]) else ""
        print("%-14s %4d  %5.1f %% samples %5.1f %% instr | %s" % (f, l, 100.0 * n / tot_s, 100.0 * ins[(f, l)] / tot_i, text))


if __name__ == "__main__":
    main()
