"""Top stall-sample lines of a kernel from `ncu -i rep --page source --csv` output (SASS view), with the stall mix.
usage: ncu_hot.py file.csv [top] [section]"""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
secs = []
for r in rows:
    if r and r[0] == "Kernel Name":
        secs.append(dict(name=r[1], rows=[]))
    elif secs:
        secs[-1]["rows"].append(r)
sec = secs[which]
print(len(secs), "sections; showing", which, sec["name"][:100])
hdr = sec["rows"][0]
body = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, " instructions executed", sum(int(r[ci["Instructions Executed"]] or 0) for r in body))
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
for i in sorted(order):
    r = body[i]
    mix = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[S]):6d} {100*int(r[S])/max(tot,1):5.1f}%  x{r[ci['Instructions Executed']]:>8}  {r[1].strip()[:70]:70s} {mix[0][1]}:{mix[0][0]} {mix[1][1]}:{mix[1][0]}")
