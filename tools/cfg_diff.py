"""Build container only: `bbc_train_cfg()` / `tsc_train_cfg()` against `class_to_dict` of the reference's config classes
(usage: python tools/cfg_diff.py bbc | tsc); prints every missing key and every differing value."""
import sys, importlib
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'quadrupedal-agility_b200'))
from ref_harness import import_reference
which=sys.argv[1]
ref=import_reference(which)
helpers=importlib.import_module("legged_gym.utils.helpers")
from qa_b200 import config as K
if which=="bbc":
    m=importlib.import_module("legged_gym.envs.go2.go2_locomotion_config")
    names=[n for n in dir(m) if n.startswith("Go2")]
    print(names)
    rc=helpers.class_to_dict(getattr(m,[n for n in names if "Algo" in n or "PPO" in n][0])())
    ours=K.bbc_train_cfg()
else:
    m=importlib.import_module("legged_gym.envs.go2.go2_agility_config")
    names=[n for n in dir(m) if n.startswith("Go2")]
    print(names)
    rc=helpers.class_to_dict(getattr(m,[n for n in names if "PPO" in n or "Algo" in n][0])())
    ours=K.tsc_train_cfg()
def walk(a,b,path=""):
    for k in sorted(set(a)|set(b)):
        if k not in b: print("  missing in ours:",path+k,"=",repr(a[k])[:80]); continue
        if k not in a: print("  extra in ours:",path+k,"=",repr(b[k])[:80]); continue
        if isinstance(a[k],dict) and isinstance(b[k],dict): walk(a[k],b[k],path+k+"."); continue
        va,vb=a[k],b[k]
        if isinstance(va,(list,tuple)): va=list(va)
        if isinstance(vb,(list,tuple)): vb=list(vb)
        if va!=vb: print(f"  DIFF {path+k}: ref={va!r} ours={vb!r}")
walk(rc,ours)
