#!/usr/bin/env python3
"""Regenerates the measurement table of README.md from the bench lines under profiles/ (r02_bench_n{1,2,4,8}.json,
r02_bench_n1_config4.json, r02_stress_config5_8gpu.json), so that the table is what was measured and nothing else.  usage: tools/readme_table.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda f: json.load(open(os.path.join(ROOT, "profiles", f)))
n1, n2, n4, n8, s5 = P("r02_bench_n1.json"), P("r02_bench_n2.json"), P("r02_bench_n4.json"), P("r02_bench_n8.json"), P("r02_stress_config5_8gpu.json")
r1, r2, r4, r8 = n1["render"], n2["render"], n4["render"], n8["render"]
r1c4 = P("r02_bench_n1_config4.json")["render"]["config4_staircase"]  # bench.py --config4 --no-cpu --no-stress, 1 GPU
c1, c3 = "config1_cornell_shell", "config3_veach_mis"
new = f'''Measured on B200 (round 2, final build; raw lines: `profiles/r02_bench_n1.json`, `r02_bench_n2.json`,
`r02_bench_n4.json`, `r02_bench_n8.json`; config 5: `r02_stress_config5_8gpu.json`, taken a few commits earlier; the table
is generated from them by `tools/readme_table.py`).  N > 1: `torchrun`, one rank per GPU for the ray batches; the
renders run inside the library (`trt_render_multi`: one process, N GPUs, scene replicated device to device, one
`ncclReduce`):

| | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs | reference on the box's 16 host cores |
|---|---|---|---|---|---|
| closest-hit, 16 Mi-ray batch per GPU vs the Cornell shell (rays resident in HBM) | {n1['value']/1e3:.1f} Grays/s | {n2['value']/1e3:.1f} | {n4['value']/1e3:.1f} | {n8['value']/1e3:.1f} Grays/s | {n1['cpu_baseline']['value']:.1f} Mrays/s |
| the same through the C ABI with host buffers (at the host-link ceiling at every N: `e2e.link_frac` {n1['e2e']['link_frac']:.2f} / {n2['e2e']['link_frac']:.2f} / {n4['e2e']['link_frac']:.2f} / {n8['e2e']['link_frac']:.2f}) | {n1['e2e']['value']/1e3:.2f} Grays/s | {n2['e2e']['value']/1e3:.2f} | {n4['e2e']['value']/1e3:.2f} | {n8['e2e']['value']/1e3:.2f} Grays/s | — |
| closest-hit, staircase (31 k triangles) / 10 M-triangle stress mesh | {n1['other_scenes']['staircase']['closest_hit_mrays']/1e3:.1f} / {n1['other_scenes']['stress_10m']['closest_hit_mrays']/1e3:.1f} Grays/s | — | — | — / {s5['closest_hit_8gpu']['mrays_per_s']/1e3:.1f} Grays/s | — |
| closest-hit, labelled stand-in for the missing cornell-box.obj (Cornell shell + 100 k-triangle sphere) | {n1['cornell_standin']['closest_hit_mrays']/1e3:.1f} Grays/s | — | — | — | — |
| config 1: Cornell shell 512x512, 16 spp | {r1[c1]['spp_per_s']:.0f} spp/s ({r1[c1]['ms']:.2f} ms) | {r2[c1]['spp_per_s']:.0f} | {r4[c1]['spp_per_s']:.0f} | {r8[c1]['spp_per_s']:.0f} spp/s ({r8[c1]['ms']:.2f} ms) | {r1[c1]['cpu_reference']['spp_per_s_at_config']:.1f} spp/s (the unmodified program, {r1[c1]['cpu_reference']['sample']}) |
| config 3: veach-mis 1280x720, 256 spp | {r1[c3]['spp_per_s']:.0f} spp/s ({r1[c3]['ms']:.0f} ms) | {r2[c3]['spp_per_s']:.0f} | {r4[c3]['spp_per_s']:.0f} | {r8[c3]['spp_per_s']:.0f} spp/s ({r8[c3]['ms']:.1f} ms) | {r1[c3]['cpu_reference']['spp_per_s_at_config']:.2f} spp/s (scaled from {r1[c3]['cpu_reference']['sample']}) |
| config 4: staircase 1920x1080, 1024 spp (37 G rays) | {r1c4['spp_per_s']:.1f} spp/s ({r1c4['ms']/1e3:.2f} s) | — | — | {r8['config4_staircase']['spp_per_s']:.0f} spp/s ({r8['config4_staircase']['ms']/1e3:.2f} s; efficiency {r1c4['ms']/(8*r8['config4_staircase']['ms']):.2f}, same 8-bit frame) | — |
| config 5: 10 M triangles, 3840x2160, 8 spp | {s5['render_1gpu_nccl']['ms']:.1f} ms | — | — | {s5['render_8gpu_nccl']['ms']:.1f} ms | — |

Round 1 → round 2 on one GPU: closest hit 15.6 → {n1['value']/1e3:.1f} Grays/s, config 1 6.46 → {r1[c1]['ms']:.2f} ms, config 3 436 → {r1[c3]['ms']:.0f} ms (staircase
1280x720x16: 104 → 82.5 ms), 10 M-triangle mesh 3.05 → {n1['other_scenes']['stress_10m']['closest_hit_mrays']/1e3:.2f} Grays/s; config 3 on 8 GPUs 57.6 → {r8[c3]['ms']:.1f} ms (efficiency 0.947 →
{r1[c3]['ms']/(8*r8[c3]['ms']):.3f}), config 4 on 8 GPUs 1.75 → {r8['config4_staircase']['ms']/1e3:.2f} s.

'''
path = os.path.join(ROOT, "README.md")
readme = open(path).read()
a = readme.index("Measured on B200 (round 2")
b = readme.index("Closest-hit triangle ids and distance bits are identical")
open(path, "w").write(readme[:a] + new + readme[b:])
print(new)
