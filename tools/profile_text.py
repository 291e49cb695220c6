#!/usr/bin/env python
"""Text summaries that go under profiles/ (run here, on files a gpurun call brought back):
  python tools/profile_text.py launches gpurun_out/launches_rNN.csv            # per-kernel launch list of an ncu
                                                                               #   --metrics gpu__time_duration.sum pass
  python tools/profile_text.py sass ir_ads_b200/libmsda_b200.so                # opcode histogram + excerpt of the two
                                                                               #   hot kernels (cuobjdump -sass)"""
import collections
import csv
import re
import subprocess
import sys

HOT = {"forward": "msda_fwd_fast_kernelILi32EfLi4ELi64ELi0ELi0ELi4E",
       "backward": "msda_bwd_fast_kernelILi32EfLi4ELi128ELi2EfLi0ELi4E"}


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        tot[r[k]] += float(r[v].replace(",", "")) / 1e6
        cnt[r[k]] += 1
    all_ms = sum(tot.values())
    print("launches   total ms  share  kernel")
    for name, ms in tot.most_common(14):
        print(f"{cnt[name]:8d} {ms:10.3f} {100 * ms / all_ms:5.1f}%  {name[:110]}")
    f = [(n, ms / cnt[n]) for n, ms in tot.items() if "msda_fwd" in n]
    b = [(n, ms / cnt[n]) for n, ms in tot.items() if "msda_bwd" in n]
    if f and b:
        fm, bm = f[0][1], b[0][1]
        print(f"\nmsda kernels only: backward {bm:.3f} ms/launch = {100 * bm / (fm + bm):.1f} % of the op's kernel time, "
              f"forward {fm:.3f} ms/launch = {100 * fm / (fm + bm):.1f} %")


def sass(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for what, key in HOT.items():
        for f in funcs[1:]:
            name = f.split("\n", 1)[0]
            if key not in name:
                continue
            lines = re.findall(r"^\s+(/\*[0-9a-f]{4}\*/\s+.*?;)", f, re.M)
            ops = collections.Counter()
            for ln in lines:
                body = ln.split("*/", 1)[1].strip()
                if body.startswith("@"):
                    body = body.split(None, 1)[1]
                ops[body.split()[0]] += 1
            print(f"== {what}: {name}\n   {len(lines)} instructions")
            for op, n in ops.most_common():
                if re.match(r"(LDG|LDS|STS|STG|RED|ATOM|FFMA|FMUL|IMAD\.WIDE|R2P|SHFL|BAR|F2I|MUFU|UTMA|UTC|LDTM|BRA|BSSY)", op):
                    print(f"   {n:5d}  {op}")
            first = next(i for i, ln in enumerate(lines) if "LDG.E.128" in ln)
            print("   excerpt (around the first LDG.E.128):")
            for ln in lines[max(0, first - 8):first + 40]:
                print("        " + ln)
            print()
            break


if __name__ == "__main__":
    {"launches": launches, "sass": sass}[sys.argv[1]](sys.argv[2])
