"""Per-source-line stall samples of an `ncu --set full --import-source on` capture (kernel compiled with -lineinfo).
  python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top]  -> lines ranked by warp-stall samples, with the dominant reasons"""
import collections, csv, io, subprocess, sys

def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    files, cur, hdr = {}, None, None
    agg = collections.defaultdict(lambda: collections.Counter())
    text = {}
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur = row[1].split("/")[-1]; hdr = None; continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row; continue
        if hdr is None or cur is None:
            continue
        if not row[0].strip():      # SASS rows under a CUDA line: already included in the line's totals
            continue
        d = dict(zip(hdr, row))
        # the same header name "Source" appears twice (CUDA line, SASS): take the first
        key = (cur, row[0])
        text.setdefault(key, row[1].strip()[:110])
        try:
            agg[key]["samples"] += int(d.get("# Samples") or 0)
            agg[key]["inst"] += int(d.get("Instructions Executed") or 0)
        except ValueError:
            continue
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                try:
                    agg[key][k] += int(v)
                except ValueError:
                    pass
    tot = sum(c["samples"] for c in agg.values()) or 1
    toti = sum(c["inst"] for c in agg.values()) or 1
    print(f"# {path}: {tot} stall samples, {toti} warp instructions")
    reasons = collections.Counter()
    for c in agg.values():
        for k, v in c.items():
            if k.startswith("stall_"):
                reasons[k] += v
    print("# by reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in reasons.most_common(10)))
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        rs = ", ".join(f"{k[6:]} {100 * v / max(c['samples'], 1):.0f}%" for k, v in c.most_common(6) if k.startswith("stall_"))
        print(f"{100 * c['samples'] / tot:5.2f}% smp {100 * c['inst'] / toti:5.2f}% inst  {key[0]}:{key[1]:>5s}  {text[key]}\n        [{rs}]")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
