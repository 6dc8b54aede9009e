"""Build container only: constructor signatures, default values and public method names of this package's classes against
the reference's (usage: python tools/api_diff.py bbc | tsc).  Found the `train_with_estimated_latent` default the TSC fork's
ActorCriticBBC flips; remaining differences are the extra acceleration switches and the reference's dead helpers."""
import sys, inspect
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'quadrupedal-agility_b200'))
from ref_harness import import_reference
which=sys.argv[1]
ref=import_reference(which)
import importlib
from qa_b200 import rsl_rl as Q
from qa_b200.rsl_rl import runner as QR, tsc_runner as QTR, algorithm as QA
pairs=[]
if which=="bbc":
    R=importlib.import_module("rsl_rl.runners.on_policy_runner")
    pairs=[(ref.actor_critic.ActorCritic,Q.ActorCritic),(ref.estimator.Estimator,Q.Estimator),(ref.discriminator.Discriminator,Q.Discriminator),
           (ref.gail.SSInfoGAIL,Q.SSInfoGAIL),(ref.RolloutStorage,Q.RolloutStorage),(R.OnPolicyRunner,QR.OnPolicyRunner),
           (ref.actor_critic.StateHistoryEncoder,Q.StateHistoryEncoder),(importlib.import_module("rsl_rl.storage.replay_buffer").ReplayBuffer,QA.ReplayBuffer),
           (ref.utils.Normalizer,Q.Normalizer)]
else:
    R=importlib.import_module("rsl_rl.runners.on_policy_runner")
    DB=importlib.import_module("rsl_rl.modules.depth_backbone")
    from qa_b200.rsl_rl import depth_backbone as QD
    pairs=[(ref.actor_critic.ActorCriticTSC,Q.ActorCriticTSC),(ref.actor_critic.Actor,Q.Actor),(ref.actor_critic.ActorCriticBBC,Q.ActorCriticBBC),
           (ref.ppo.PPO,Q.PPO),(ref.RolloutStorage,Q.RolloutStorageTSC),(ref.discriminator.Discriminator,Q.DiscriminatorTSC),
           (R.OnPolicyRunner,QTR.OnPolicyRunnerTSC),(DB.DepthOnlyFCBackbone58x87,QD.DepthOnlyFCBackbone58x87),(DB.RecurrentDepthBackbone,QD.RecurrentDepthBackbone),
           (ref.estimator.Estimator,Q.Estimator)]
for a,b in pairs:
    sa,sb=inspect.signature(a.__init__),inspect.signature(b.__init__)
    pa,pb=sa.parameters,sb.parameters
    print("==",a.__name__,"->",b.__name__)
    la=[k for k in pa if k not in("self",)]; lb=[k for k in pb if k!="self"]
    pos_a=[k for k in la if pa[k].default is inspect._empty and pa[k].kind in (1,)]
    pos_b=[k for k in lb if pb[k].default is inspect._empty and pb[k].kind in (1,)]
    if pos_a!=pos_b: print("  positional differ:\n   ref ",pos_a,"\n   ours",pos_b)
    for k in la:
        if k in pb:
            da,db=pa[k].default,pb[k].default
            if da is not inspect._empty and da!=db: print(f"  default differs {k}: ref={da!r} ours={db!r}")
        elif pa[k].kind not in (2,4): print("  missing in ours:",k, "(default",pa[k].default,")")
    for k in lb:
        if k not in pa and pb[k].kind not in (2,4): print("  extra in ours:",k)
    # methods
    ma={n for n,_ in inspect.getmembers(a, predicate=inspect.isfunction) if not n.startswith('_')}
    mb={n for n in dir(b) if not n.startswith('_')}
    miss=sorted(ma-mb)
    if miss: print("  methods missing in ours:",miss)
