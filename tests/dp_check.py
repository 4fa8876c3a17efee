"""Run under torchrun (N ranks, one GPU each): data-parallel FusedTrainer on a sharded batch must equal the
single-process FusedTrainer on the global batch (BatchNorm statistics and gradients all-reduced).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecg_denoise_b200 import synth  # noqa: E402
from ecg_denoise_b200.engine import FusedTrainer  # noqa: E402
from ecg_denoise_b200.model import transformer  # noqa: E402
from oracle import synth_weights as SW  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    B = 8 * world
    noisy, clean = synth.make_batch(B, 2, 256, seed=123)
    sd = SW.make_state_dict("rw", 1, 5)
    use_graph = os.environ.get("DP_GRAPH", "0") == "1"

    def run(x, t, pg_world):
        m = transformer.ralenet(high_level_enhence=True)
        m.load_state_dict(sd)
        m = m.to(dev)
        tr = FusedTrainer(m, lr=1e-3, use_graph=use_graph)
        if pg_world == 1:
            tr.world = 1
        losses = []
        for _ in range(3):
            loss, _, _, _ = tr.step(x, t)
            losses.append(loss.item())
        return m, losses

    lo, hi = rank * B // world, (rank + 1) * B // world
    m_dp, l_dp = run(torch.from_numpy(noisy[lo:hi]).to(dev), torch.from_numpy(clean[lo:hi]).to(dev), world)
    ok = True
    if rank == 0:
        m_1, l_1 = run(torch.from_numpy(noisy).to(dev), torch.from_numpy(clean).to(dev), 1)
        worst = 0.0
        for (n, p), (_, q) in zip(m_dp.named_parameters(), m_1.named_parameters()):
            a, b = p.detach().double().cpu().numpy(), q.detach().double().cpu().numpy()
            if n.endswith("to_kv.bias"):
                # key half: exact gradient 0 (softmax shift invariance), Adam normalises rounding noise -> not
                # comparable; the value half is compared (see tests/test_gpu_net.py)
                a, b = a[a.size // 2:], b[b.size // 2:]
            e = np.max(np.abs(a - b) / (np.abs(b) + np.sqrt((b * b).mean()) + 1e-30))
            worst = max(worst, e)
        bn_d, bn_1 = m_dp.conv1[2], m_1.conv1[2]
        e_bn = float((bn_d.running_var - bn_1.running_var).abs().max())
        ok = worst < 1e-3 and np.allclose(l_dp, l_1, rtol=1e-4) and e_bn < 1e-5
        comm = getattr(m_dp._plan, "_comm", None)
        print(f"dp_check exchange path: {comm.kind if comm is not None else 'NCCL all-reduce'}")
        print(f"dp_check world={world} graph={use_graph}: losses dp {l_dp} single {l_1}; worst param rel err {worst:.2e}; "
              f"bn running_var diff {e_bn:.2e} -> {'OK' if ok else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
