"""Host-side multi-GPU plumbing: games are independent, so ranks own disjoint ranges of global game ids
and the only collective is one end-of-run reduction of timing / episode statistics (SURVEY.md §8 e)."""
from dataclasses import dataclass


def shard_range(batch_index: int, world: int, rank: int, games_per_rank: int):
    """Global game ids [lo, hi) owned by `rank` for batch `batch_index` (weak scaling: fixed per-rank batch).
    Seeds equal global ids, so results are invariant to the number of ranks."""
    lo = (batch_index * world + rank) * games_per_rank
    return lo, lo + games_per_rank


def shard_paths(paths, world: int, rank: int):
    """Replay ingestion over several GPUs: logs are independent, so rank r reads every world-th file starting at r
    (`ReplayBatch.from_files(shard_paths(paths, world, rank))` per rank, no exchange on the data path — what the reference's
    datasets do per DataLoader worker, riichienv-ml/.../datasets/mjai_logs.py:66-69)."""
    return list(paths)[rank::world]


@dataclass
class RunStats:
    elapsed_ms: float = 0.0       # device time of this rank's timed region
    kernel_ms: float = 0.0
    e2e_s: float = 0.0
    env_steps: float = 0.0
    e2e_steps: float = 0.0
    games: float = 0.0
    score_sum: float = 0.0        # sum of final scores over all seats (conservation check: 100000 per game - sticks)

    FIELDS = ("elapsed_ms", "kernel_ms", "e2e_s", "env_steps", "e2e_steps", "games", "score_sum")

    def tensor(self, torch, device):
        return torch.tensor([getattr(self, f) for f in self.FIELDS], dtype=torch.float64, device=device)


def reduce_stats(stats: RunStats, dist, torch, device) -> RunStats:
    """MAX over ranks for times, SUM for counts.  `dist` = torch.distributed (nccl on GPUs, gloo in CPU tests) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return stats
    t = stats.tensor(torch, device)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    out = RunStats()
    for i, f in enumerate(RunStats.FIELDS):
        setattr(out, f, (mx if f in ("elapsed_ms", "kernel_ms", "e2e_s") else sm)[i].item())
    return out
