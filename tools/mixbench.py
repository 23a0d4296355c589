import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tsadar_b200.engine import microbench
print("ffma", microbench(0)/1e12, "lg2", microbench(1)/1e12)
for nf in [0, 2, 4, 6, 8, 10, 12]:
    r = microbench(100 + nf, 2048)
    # per MUFU: nf FFMA + 1 FADD
    clk = 148 * 1.965e9 * 128 / 32 / r   # SMSP-cycles per warp-level MUFU
    print(f"NF={nf:2d} FFMA per MUFU: {r/1e12:.3f} T MUFU/s  -> {clk:.2f} clk per (MUFU + {nf} FFMA + 1 FADD) warp-group")
print("ffma2 peak (T FMA/s)", microbench(3)/1e12)
for nf2 in [1, 2, 3, 4, 5, 6]:
    r = microbench(200 + nf2, 2048)
    print(f"NF2={nf2} FFMA2 (= {2*nf2} FMA) per MUFU: {r/1e12:.3f} T MUFU/s -> {3.722e13/r:.2f} SMSP-clk per warp-MUFU")
