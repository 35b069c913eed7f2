"""Map the tc_kernel launches of one train step (ncu launch list, gpu__time_duration) to U-Net-CA layers and print
per-layer algorithmic TFLOP/s.  usage: python profiles/layer_table.py launches.csv [B] [S]"""
import csv, sys
W = (64, 128, 256, 512, 1024)

def expected_calls(B, S, cin=3):
    calls = []
    def conv(kind, C, O, l): calls.append((f"{kind} {C}->{O} @{S>>l}", 2.0*B*(S>>l)**2*9*C*O))
    def convT(kind, Cin, Cout, l): calls.append((f"{kind} {Cin}->{Cout} @{S>>l}", 2.0*B*(S>>l)**2*Cin*4*Cout))
    # forward
    calls.append((f"first_fwd {cin}->64 @{S}", 2.0*B*S*S*9*cin*64)); conv("fwd", 64, 64, 0)
    for l in range(1, 5): conv("fwd", W[l-1], W[l], l); conv("fwd", W[l], W[l], l)
    for l in (3, 2, 1, 0):
        convT("convT_fwd", 2*W[l], W[l], l+1); conv("fwd", 2*W[l], W[l], l); conv("fwd", W[l], W[l], l)
    # backward
    for l in range(4):
        conv("wgrad", W[l], W[l], l); conv("dgrad", W[l], W[l], l); conv("wgrad", 2*W[l], W[l], l); conv("dgrad", W[l], 2*W[l], l)
        convT("convT_wgrad", 2*W[l], W[l], l+1); convT("convT_dgrad", 2*W[l], W[l], l+1)
    for l in range(4, 0, -1):
        conv("wgrad", W[l], W[l], l); conv("dgrad", W[l], W[l], l); conv("wgrad", W[l-1], W[l], l); conv("dgrad", W[l], W[l-1], l)
    conv("wgrad", 64, 64, 0); conv("dgrad", 64, 64, 0); calls.append((f"first_wgrad {cin}->64 @{S}", 2.0*B*S*S*9*cin*64))
    return calls

def main():
    path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64; S = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    def ms(r):
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        return v/1e3 if u == "us" else v/1e6 if u == "ns" else v*1e3 if u == "s" else v
    # find a step start: the im2col kernel
    starts = [i for i, r in enumerate(rows) if "im2col3x3" in r["Kernel Name"]]
    i0 = starts[0]; i1 = starts[1] if len(starts) > 1 else len(rows)
    step = rows[i0:i1]
    tc = [(r["Kernel Name"].split("(")[0].replace("void unetca::", ""), ms(r)) for r in step if "tc_kernel" in r["Kernel Name"] or "tc_wgrad3x3" in r["Kernel Name"]]
    exp = expected_calls(B, S)
    print(f"{len(tc)} tc launches in the step, {len(exp)} expected; step total {sum(ms(r) for r in step):.1f} ms (serialised, cold)")
    tot_ms = tot_fl = 0
    for (name, fl), (k, t) in zip(exp, tc):
        tot_ms += t; tot_fl += fl
        print(f"{name:28s} {k:22s} {t:8.3f} ms {fl/t/1e9:8.1f} TFLOP/s")
    print(f"tensor total {tot_ms:.1f} ms, {tot_fl/tot_ms/1e9:.1f} TFLOP/s")
main()
