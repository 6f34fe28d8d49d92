"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with the line table of the
matching cubin (nvdisasm -g -c) and prints, per kernel, executed warp instructions and stall samples by source line.
usage: python tools/ncu_lines.py sass.csv disasm.txt kernel_substring [top_n]"""
import collections
import csv
import re
import sys


def load_disasm(path):
    """kernel (mangled) -> list of (opcode text, file, line, inline chain)"""
    out = collections.defaultdict(list)
    fn, cur = None, ("?", 0, "")
    for line in open(path):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "[^"]*/([^"/]*)", line (\d+)(.*)', line)
        if m:
            inl = re.findall(r'inlined at "[^"]*/([^"/]*)", line (\d+)', m.group(3))
            cur = (m.group(1), int(m.group(2)), " <- ".join("%s:%s" % x for x in inl))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and fn:
            out[fn].append((m.group(2).strip(), cur))
    return out


def main():
    sass_csv, disasm, want = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = load_disasm(disasm)
    rows = list(csv.reader(open(sass_csv)))
    i = 0
    seen = set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(dict(zip(hdr, rows[j])))
                j += 1
            i = j
            if want not in name or name in seen:
                continue
            seen.add(name)
            key = [k for k in dis if want in k and len(dis[k]) == len(body)]
            if not key:
                print("no disassembly of matching length for", name, len(body), [(k[:40], len(dis[k])) for k in dis if want in k])
                continue
            d = dis[key[0]]
            by_line = collections.Counter()
            smp_line = collections.Counter()
            by_fn = collections.Counter()
            tot = 0
            for b, (op, (f, l, inl)) in zip(body, d):
                n = float(b["Instructions Executed"] or 0)
                s = float(b["# Samples"] or 0)
                by_line[(f, l)] += n
                smp_line[(f, l)] += s
                tot += n
            ts = sum(smp_line.values())
            print("== %s: %.3g warp instructions, %d samples" % (name[:60], tot, ts))
            for (f, l), n in by_line.most_common(top):
                print("  %5.1f%% inst  %5.1f%% samples   %s:%d" % (100 * n / tot, 100 * smp_line[(f, l)] / max(ts, 1), f, l))
        else:
            i += 1


if __name__ == "__main__":
    main()
