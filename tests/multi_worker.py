"""Worker for tests/test_multi_gloo.py: the N>1 host logic on the gloo backend (CPU).
The per-rank "system build" is the CPU oracle here (tests may use it); on the GPUs it is
ssf_icp_build -- the sharding, the ordered reduction and the loop are the same code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleShardEngine:
    """Duck-types the step-wise ICP methods of SupersurfelFusion with the oracle's arithmetic."""

    def __init__(self, orc, cam, prob, icp_iter=10):
        import types
        self.orc, self.cam, self.p = orc, cam, prob
        self.cfg = types.SimpleNamespace(icp_iter=icp_iter)

    def icpBegin(self, R_init, t_init):
        self.R0 = np.asarray(R_init, np.float32).reshape(3, 3)
        self.t0 = np.asarray(t_init, np.float32).reshape(3)
        self.tf = np.eye(4)
        self.systems = []

    def _transform(self):
        Ri = self.tf[:3, :3].astype(np.float32)
        ti = self.tf[:3, 3].astype(np.float32)
        return (Ri @ self.R0).astype(np.float32), (Ri @ self.t0 + ti).astype(np.float32)

    def icpBuild(self, begin, count):
        p = self.p
        R, t = self._transform()
        sl = slice(begin, begin + count)
        return self.orc.icp_system(self.cam, p["src_pos"][sl], p["src_col"][sl], p["src_ori"][sl], p["tgt_col"],
                                   p["tgt_ori"], p["tgt_conf"], R, t, p["labels"], p["depth"])

    def icpSolve(self, s):
        # one Gauss-Newton step in numpy (enough for the test: 2 iterations, no convergence logic)
        A = np.zeros((6, 6))
        k = 0
        for i in range(6):
            for j in range(i, 6):
                A[i, j] = A[j, i] = s[k]
                k += 1
        x = np.linalg.solve(A, s[21:27].astype(np.float64))
        self.systems.append(np.array(s))
        ax, tr = x[:3], x[3:]
        n = np.linalg.norm(ax)
        ang = 0.5 * np.arctan(n)
        ax = ax / n
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rr = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        T = np.eye(4)
        T[:3, :3] = Rr @ Rr
        T[:3, 3] = Rr @ (tr * np.cos(ang))
        self.tf = T @ self.tf
        return len(self.systems) >= 2

    def icpFinish(self, apply_to_pose):
        return True, self.tf[:3, :3].T.astype(np.float32), np.zeros(3, np.float32), dict(iters=len(self.systems), valid=1)


def run(rank, world, port, outdir):
    import torch.distributed as dist
    from oracle import orc
    from supersurfel_fusion_b200 import multi
    from supersurfel_fusion_b200.synth import synthetic_icp_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = synthetic_icp_problem(20011, width=320, height=240, seed=3)
    cam = orc.OrcCam(*prob["cam"])
    eng = OracleShardEngine(orc, cam, prob)
    R = np.eye(3, dtype=np.float32)
    t = np.array([0.002, -0.001, 0.003], np.float32)
    valid, Rr, tr, info = multi.tile_parallel_icp(eng, dist, len(prob["src_pos"]), R, t)
    # timing logic of bench.py: max over ranks, whole-job aggregate
    ms = 10.0 + rank
    mx = multi.allreduce_max(dist, ms)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), systems=np.array(eng.systems), shard=np.array(info["shard"]),
             max_ms=mx, fps=multi.aggregate_throughput(100, world, mx), seed=multi.sequence_seed(rank), Rr=Rr)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
