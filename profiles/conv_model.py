"""Closed-form time model of the conv GEMMs of one batch-256 ResNet-50 step, built from the round-1 measurements
(profiles/r1_conv_layers.md): every 128 x BN tile re-fetches its operands per 64-deep k-block through L2, so a launch
costs about

    waves x max(k-blocks x t_kb, t_epilogue) + t_fixed,     t_kb = max(bytes_per_kblock x SMs / L2_BW, t_gather, t_mma)

with L2_BW = 12.4 TB/s (the chip's L2->SM cap), t_gather = 0.55 us per 16 KB gathered A tile, t_mma = the tensor floor
(128 x BN x 64 MACs at 4096 MAC/clk/SM), t_epilogue = 2.4 us per 128 x 256 tile (scaled by BN), t_fixed = 5 us, and
never less than the DRAM floor.  Prints the modelled time per GEMM class for (a) the shipped kernels, (b) CTA pairs
with both operands by TMA (32 KB per k-block at BN = 256), so that round 2 can rank its options.

    python profiles/conv_model.py > profiles/r1_conv_model.md
"""
SMS, CLK = 148, 1.965e9
L2_BW, DRAM_BW = 12.4e12, 6.561e12
T_GATHER, T_FIXED, T_EPI256 = 0.55e-6, 5e-6, 2.4e-6


def convs():
    out = [("stem 7x7/2", 224, 3, 64, 7, 2, 3)]
    h, inpl = 56, 64
    for li, nb in enumerate((3, 4, 6, 3)):
        pl = 64 << li
        for b in range(nb):
            s = 2 if (b == 0 and li > 0) else 1
            out += [(f"l{li+1}.{b}.c1", h, inpl, pl, 1, 1, 0), (f"l{li+1}.{b}.c2", h, pl, pl, 3, s, 1),
                    (f"l{li+1}.{b}.c3", h // s, pl, pl * 4, 1, 1, 0)]
            if b == 0:
                out.append((f"l{li+1}.{b}.ds", h, inpl, pl * 4, 1, s, 0))
            inpl, h = pl * 4, h // s
    return out


def pick_bn(n, m_tiles, gather):
    if n % 256 == 0:
        t = m_tiles * (n // 256)
        if t >= 2 * SMS or (gather and 2 * t >= SMS):
            return 256
    return 128 if n % 128 == 0 else 64


def gemm_time(m, n, k_blocks, gather, dram_bytes, pairs=False):
    """One fprop / dgrad style GEMM: M rows, N columns, k_blocks of 64."""
    m_tiles = -(-m // 128)
    bn = pick_bn(n, m_tiles, gather)
    tiles = m_tiles * (n // bn)
    paired = pairs and bn >= 128
    b_rows = bn // 2 if paired else bn
    bytes_kb = (128 + b_rows) * 128
    t_mma = 128 * bn * 64 / 4096 / CLK
    t_kb = max(bytes_kb * SMS / L2_BW, t_mma, T_GATHER if (gather and not paired) else 0.0)
    waves = -(-tiles // SMS)
    t_tile = max(k_blocks * t_kb, T_EPI256 * bn / 256)
    return max(waves * t_tile + T_FIXED, dram_bytes / DRAM_BW), bn


def main():
    n = 256
    tot = {"shipped": [0.0, 0.0], "pairs": [0.0, 0.0]}
    rows = []
    for name, h, cin, cout, k, s, p in convs():
        if name.startswith("stem"):
            continue
        ho = (h + 2 * p - k) // s + 1
        m_out, m_in = n * ho * ho, n * h * h
        kb_f, kb_d = k * k * cin // 64, k * k * cout // 64
        plain = k == 1 and s == 1
        dram = (m_in * cin + m_out * cout) * 2
        for key, pairs in (("shipped", False), ("pairs", True)):
            tf, bnf = gemm_time(m_out, cout, kb_f, not plain, dram, pairs)
            if s == 1:
                td, _ = gemm_time(m_in, cin, kb_d, not plain, dram, pairs)
            else:       # four parity classes, each a quarter of the rows and of the taps
                td = sum(gemm_time(m_in // 4, cin, max(1, kb_d * t // 9 if k == 3 else kb_d), True, dram / 4, False)[0]
                         for t in ((1, 2, 2, 4) if k == 3 else (1,)))
            tot[key][0] += tf
            tot[key][1] += td
            if key == "shipped":
                rows.append([name, f"{cin}->{cout} {k}x{k}/{s} @{h}", bnf, tf * 1e6, td * 1e6])
            else:
                rows[-1] += [tf * 1e6, td * 1e6]
    print("# Modelled conv GEMM times, batch 256 (profiles/conv_model.py; constants from the round-1 measurements)\n")
    print("| conv | shape | BN | fprop us (shipped) | dgrad us (shipped) | fprop us (TMA-fed pairs) | dgrad us (TMA-fed pairs) |")
    print("|---|---|---:|---:|---:|---:|---:|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.0f} | {r[4]:.0f} | {r[5]:.0f} | {r[6]:.0f} |")
    print(f"\n| total over the 52 non-stem convs | fprop ms | dgrad ms |\n|---|---:|---:|")
    print(f"| shipped kernels (model) | {tot['shipped'][0]*1e3:.2f} | {tot['shipped'][1]*1e3:.2f} |")
    print(f"| measured in the step (bench.py breakdown, incl. the stem's 0.32 ms fprop) | 4.09 | 3.99 |")
    print(f"| CTA pairs, both operands by TMA (model) | {tot['pairs'][0]*1e3:.2f} | {tot['pairs'][1]*1e3:.2f} |")


if __name__ == "__main__":
    main()
