"""torchrun helper of tests/test_gpu_api.py::test_two_rank_ppo_step_keeps_parameters_identical and of bench.py's SCALE proof:
each rank steps its own env shard (factory_marl_b200.dist.env_shard), one collect_rollouts + train with the NCCL gradient
all-reduce, then the parameter checksums of all ranks are compared."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from factory_marl_b200 import dist as fdist  # noqa: E402
from factory_marl_b200.ppo import PPO  # noqa: E402
from factory_marl_b200.vec_env import FactoryVecEnv  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
total = 512
lo, hi = fdist.env_shard(total, rank, world)
kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=hi - lo, num_arms=2, seeds=fdist.shard_seeds(42, total, rank, world),
                    device=f"cuda:{torch.cuda.current_device()}", **kw)
ppo = PPO(env, n_steps=4, batch_size=256, n_epochs=2, seed=0)
for _ in range(2):
    ppo.collect_rollouts(); ppo.train()
chk = torch.stack([ppo.pol.params.double().sum(), ppo.pol.params.double().abs().sum()])
all_chk = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(all_chk, chk)
if rank == 0:
    print("PARAMS_EQUAL", all(torch.equal(all_chk[0], c) for c in all_chk), [c.tolist() for c in all_chk])
dist.destroy_process_group()
