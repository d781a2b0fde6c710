"""Fills the @...@ placeholders of profiles/README.md's round-2 tables from a bench.py JSON line (and the reference arm's)."""
import json, sys, re
bench, ref = sys.argv[1], sys.argv[2]
d = json.loads([l for l in open(bench) if l.startswith('{')][-1])
r = json.loads([l for l in open(ref) if l.startswith('{')][-1])
ro, r2 = d['roofline'], d['roofline_1280x960']
lm = d['large_map']
m5, m20 = lm['640x480-5M'], lm['1280x960-20M']
def st(m, k):
    v = m['map_stage_rooflines'][k]
    return f"{v['duration_us']:.0f} µs, {v['achieved']:.0f} GB/s ({v['frac']:.2f})"
sub = {
    'V640': f"{d['value']:.0f}", 'MS640': f"{d['ms_per_step']:.3f}", 'E640': f"{d['e2e']['value']:.0f}",
    'V1280': f"{d['value_1280x960']['value']:.0f} ({d['value_1280x960']['ms_per_step']:.2f} ms)", 'E1280': f"{d['value_1280x960']['e2e']:.0f}",
    'V5M': f"{m5['value']:.0f} ({m5['ms_per_step']:.2f} ms)", 'E5M': f"{m5['e2e']:.0f}",
    'V20M': f"{m20['value']:.0f} ({m20['ms_per_step']:.2f} ms)", 'E20M': f"{m20['e2e']:.0f}",
    'NV': f"{d['no_lookahead']['value']:.0f}", 'NE': f"{d['no_lookahead']['e2e']:.0f}",
    'REF': f"{r['value']:.0f} frames/s over its first 40 frames (`--impl reference`); {d['cpu_baseline']['value']:.1f} over 170 frames (`cpu_baseline`)",
    'TO': f"{d['tracking_only']['ours_ms']:.3f}", 'TR': f"{d['tracking_only']['reference_ms']:.1f}",
    'LPF': f"{d['launches_per_frame']:.0f}",
    'IDX5': st(m5, 'index_map'), 'IDX20': st(m20, 'index_map'), 'RAY5': st(m5, 'raycast'), 'RAY20': st(m20, 'raycast'),
    'CS5': st(m5, 'clean_static'), 'CS20': st(m20, 'clean_static'), 'CM5': st(m5, 'clean_shift'), 'CM20': st(m20, 'clean_shift'),
    'R640': f"{ro['duration_us']:.2f}", 'F640': f"{ro['frac']:.2f}", 'R1280': f"{r2['duration_us']:.1f}", 'F1280': f"{r2['frac']:.2f}",
    'W640': f"{ro['duration_warm_us']:.2f}", 'W1280': f"{r2['duration_warm_us']:.1f}",
    'FI640': f"{ro['full_iteration']['duration_us']:.1f}", 'FI1280': f"{r2['full_iteration']['duration_us']:.1f}",
    'FS': '2.5',
}
p = 'profiles/README.md'
s = open(p).read()
for k, v in sub.items():
    s = s.replace('@' + k + '@', v)
left = re.findall(r'@[A-Z0-9]+@', s)
open(p, 'w').write(s)
print('left:', left)
