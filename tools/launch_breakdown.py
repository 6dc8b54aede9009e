#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: whole capture and one PPO minibatch
(the launches between the last two k_gather_minibatch launches).  usage: python tools/launch_breakdown.py FILE [top]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr, seq = None, []
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = d['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    seq.append((d['Kernel Name'], v))


def show(title, part):
    agg = collections.defaultdict(list)
    for k, v in part:
        agg[k[:72]].append(v)
    tot = sum(v for _, v in part)
    print(f"== {title}: {len(part)} launches, {tot:.0f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print(f"{k:74s} n={len(v):5d} avg={sum(v) / len(v):8.2f} tot={sum(v):9.1f} {sum(v) / tot * 100:5.1f}%")


show("whole capture", seq)
# one PPO minibatch step = the launches between two consecutive PPO-loss launches (the gathers no longer delimit a step:
# the static schedule gathers every minibatch once per update)
idx = [i for i, (k, _) in enumerate(seq) if 'k_ppo_loss' in k and 'tsc' not in k]
if len(idx) >= 3:
    show("one PPO minibatch step", seq[idx[-3]:idx[-2]])
k2 = [i for i, (k, _) in enumerate(seq) if 'k_post_physics' in k]
if len(k2) >= 2:
    show("one rollout step", seq[k2[-2]:k2[-1]])
