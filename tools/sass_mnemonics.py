"""SASS mnemonics of the tensor-core kernels in libsurj_b200.so (cuobjdump -sass), counted per kernel -> profiles/."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "surjectors_b200", "_lib", "libsurj_b200.so")
KEEP = re.compile(r"^(UTC\w+|LDTM|STTM|UBLKCP|UTMALDG|UTMAPF|LDGSTS|SYNCS)")


def main(out):
  sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
  names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
  per, order, cur, i = {}, [], None, 0
  for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = names[i][:110]
      i += 1
      per[cur] = [0, collections.Counter()]
      order.append(cur)
      continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur:
      per[cur][0] += 1
      op = m.group(1)
      if KEEP.match(op):
        per[cur][1][op] += 1
  lines = ["# SASS mnemonics of the tensor-core kernels in libsurj_b200.so (cuobjdump -sass, sm_100a), counted per kernel (tools/sass_mnemonics.py).",
           "# UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA tensor load), UBLKCP = cp.async.bulk,",
           "# LDGSTS = cp.async, UTCBAR = tcgen05.commit, SYNCS.* = mbarrier", ""]
  for k in sorted(order):
    n, c = per[k]
    if not any(o.startswith("UTC") and "MMA" in o for o in c):
      continue
    lines.append(k)
    lines.append(f"   instructions {n}; " + ", ".join(f"{o} x{v}" for o, v in sorted(c.items())))
  open(out, "w").write("\n".join(lines) + "\n")
  print("\n".join(lines))


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_mnemonics.txt"))
