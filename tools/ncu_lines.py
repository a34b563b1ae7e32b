"""Per-source-line instruction / stall-sample shares and the SASS opcode mix of every kernel in an .ncu-rep captured with
--import-source on (runs where ncu is installed; no GPU needed):  python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kern, cur, hdr = None, None, None
lines, ops, seen = {}, {}, set()
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "Function Name":
        kern = r[1][:70]; lines.setdefault(kern, {}); ops.setdefault(kern, collections.Counter()); continue
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10 or kern is None:
        continue
    try:
        n, s = int(r[7]), int(r[6])
    except ValueError:
        continue
    if r[0].isdigit() and r[2] == "-":
        a = lines[kern].setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:100]])
        a[0] += n; a[1] += s
    elif r[2].startswith("0x") and (kern, r[2]) not in seen:
        seen.add((kern, r[2]))
        t = r[3].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[kern][op.split(".")[0]] += n
for k in lines:
    tot = sum(a[0] for a in lines[k].values()) or 1
    ts = sum(a[1] for a in lines[k].values()) or 1
    print(f"\n=== {k}: {sum(ops[k].values())} warp instructions, {ts} stall samples")
    print("   opcodes: " + ", ".join(f"{o} {100 * v / max(1, sum(ops[k].values())):.1f}%" for o, v in ops[k].most_common(12)))
    for (f, l), (n, s, src) in sorted(lines[k].items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / ts))[:top]:
        print(f"  instr {100 * n / tot:5.1f}%  stall {100 * s / ts:5.1f}%  {f}:{l}: {src}")
