"""Rewrites the number-carrying rows of profiles/README.md's round-2 tables from a bench.py JSON line and the reference arm's
(idempotent: rows are found by their fixed first cell).  python scripts/fill_profiles_readme.py <bench.json> <reference.json>"""
import json, re, sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
r = json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1])
p = 'profiles/README.md'
s = open(p).read()
lm = d['large_map']
m5, m20 = lm['640x480-5M'], lm['1280x960-20M']


def st(m, k):
    v = m['map_stage_rooflines'][k]
    return f"{v['duration_us']:.0f} µs, {v['achieved']:.0f} GB/s ({v['frac']:.2f})"


def row(prefix, new):
    global s
    i = s.index(prefix)
    j = s.index('\n', i)
    s = s[:i] + new + s[j:]


ro, r2 = d['roofline'], d['roofline_1280x960']
row("| `value` (frames resident in HBM", f"| `value` (frames resident in HBM, device time, L2 flushed between frames) | **{d['value']:.0f} frames/s** ({d['ms_per_step']:.3f} ms) | {d['value_1280x960']['value']:.0f} ({d['value_1280x960']['ms_per_step']:.2f} ms) | {m5['value']:.0f} ({m5['ms_per_step']:.2f} ms) | {m20['value']:.0f} ({m20['ms_per_step']:.2f} ms) |")
row("| `e2e` (host buffers", f"| `e2e` (host buffers → pinned → H2D → frame → pose read back, every frame) | **{d['e2e']['value']:.0f}** | {d['value_1280x960']['e2e']:.0f} | {m5['e2e']:.0f} | {m20['e2e']:.0f} |")
row("| same, plain calls", f"| same, plain calls (no look-ahead: frame i+1 not available while frame i runs) | {d['no_lookahead']['value']:.0f} / {d['no_lookahead']['e2e']:.0f} | | | |")
row("| reference arm on the same box", f"| reference arm on the same box (reference CUDA tracking kernels + CPU-oracle mapping, 16 host threads per sequence) | {r['value']:.0f} frames/s over its first 40 frames (`--impl reference`); {d['cpu_baseline']['value']:.1f} over 170 frames (`cpu_baseline`) | | | |")
row("| tracking only", f"| tracking only (init* + getIncrementalTransformation): ours vs the reference's own kernels and host loop | {d['tracking_only']['ours_ms']:.3f} ms vs {d['tracking_only']['reference_ms']:.1f} ms | | | |")
row("| launches per frame", f"| launches per frame | {d['launches_per_frame']:.0f} (was 85) | | | |")
row("| index map (`k_index_scatter`", f"| index map (`k_index_scatter` + `k_index_resolve`), whole map | {st(m5, 'index_map')} | {st(m20, 'index_map')} | 32 B / surfel + 52 B / px |")
row("| model raycast (`k_splat_scatter`", f"| model raycast (`k_splat_scatter` + `k_splat_resolve`) | {st(m5, 'raycast')} | {st(m20, 'raycast')} | 32 B / surfel + 38 B / px |")
row("| clean, nothing moves", f"| clean, nothing moves (`k_clean_flags` + `k_clean_move`) | {st(m5, 'clean_static')} (fused single kernel of mid-round: 267 µs) | {st(m20, 'clean_static')} (1239 µs) | 32 B / surfel (+16 B for those in view) |")
row("| clean, every surfel moves", f"| clean, every surfel moves down by one | {st(m5, 'clean_shift')} (310 µs) | {st(m20, 'clean_shift')} (1781 µs) | 32 + 48 + 48 B / surfel |")
row("| `bench.py` roofline: dense pass, L2-cold", f"| `bench.py` roofline: dense pass, L2-cold, average launch | {ro['duration_us']:.2f} µs → `frac` {ro['frac']:.2f} | {r2['duration_us']:.1f} µs → `frac` {r2['frac']:.2f} |")
row("| same, L2-warm |", f"| same, L2-warm | {ro['duration_warm_us']:.2f} µs | {r2['duration_warm_us']:.1f} µs |")
row("| complete iteration `k_iter1` + `k_iter2`", f"| complete iteration `k_iter1` + `k_iter2` (dense rows → partials → double sums → 6×6 LDLᵀ → pose), L2-cold | {ro['full_iteration']['duration_us']:.1f} µs | {r2['full_iteration']['duration_us']:.1f} µs |")
s = re.sub(r"so the round went 1318 → \d+\.", f"so the round went 1318 → {d['value']:.0f}.", s)
open(p, 'w').write(s)
print('value', d['value'])
