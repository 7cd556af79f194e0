"""Aggregate the warp-stall samples of one kernel in an ncu report by CUDA source line.
The report's SASS listing (--page source) is zipped with `nvdisasm --print-line-info` of the same kernel in the
shipped library (same instruction order).  Usage:
  python tools/ncu_lines.py <report.ncu-rep> <kernel regex for ncu> <object basename> <mangled-name substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, obj, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several launches may match: keep the first kernel block
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = blocks[1] if len(blocks) > 1 else len(rows)
hdr, data = rows[blocks[0] + 1], rows[blocks[0] + 2:end]
ins = hdr.index("# Samples")
keys = ["stall_long_sb", "stall_barrier", "stall_short_sb", "stall_mio", "stall_wait", "stall_math", "stall_lg",
        "stall_selected", "stall_not_selected"]
cols = {k: hdr.index(k) for k in keys}
with tempfile.TemporaryDirectory() as t:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "sp-nerf_b200/lib/libspnerf_sm100a.so")], cwd=t,
                   capture_output=True)
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(t, obj + ".sm_100a.cubin")],
                         capture_output=True, text=True).stdout
fn = None; line = None; lst = []
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if fn and sub in fn and re.match(r"\s*/\*[0-9a-f]+\*/\s+.*?;", l): lst.append(line)
if len(lst) != len(data):
    sys.exit("instruction counts differ: library %d, report %d (library rebuilt since the capture?)" % (len(lst), len(data)))
agg = collections.defaultdict(collections.Counter)
for line, r in zip(lst, data):
    agg[line]["n"] += int(r[ins] or 0)
    for k, i in cols.items(): agg[line][k] += int(r[i] or 0)
tot = sum(v["n"] for v in agg.values())
print("total samples", tot)
for line, v in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    print("%22s:%-4d %6d %5.1f%%  long %5d bar %5d short %5d mio %5d wait %5d math %4d lg %4d sel %5d" % (
        line[0], line[1], v["n"], 100.0 * v["n"] / tot, v["stall_long_sb"], v["stall_barrier"], v["stall_short_sb"],
        v["stall_mio"], v["stall_wait"], v["stall_math"], v["stall_lg"], v["stall_selected"] + v["stall_not_selected"]))
