"""Dynamic opcode histogram and per-function instruction counts from an ncu report captured with --import-source on."""
import collections
import csv
import re
import subprocess
import sys


def main(rep, src_file=None):
  txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
  rows = list(csv.reader(txt.splitlines()))
  hdr = None
  ops = collections.Counter()
  tot = 0
  for r in rows:
    if len(r) > 2 and r[0] == "Address":
      hdr = r
      continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
      ie = int(r[hdr.index("Instructions Executed")])
      m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
      if m:
        ops[m.group(1)] += ie
        tot += ie
  print("warp instructions:", tot)
  print(", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(24)))


if __name__ == "__main__":
  main(*sys.argv[1:])
