"""Times the reference FORMULATION (oracle/, plain functional PyTorch, un-collapsed head) on the GPU: the
"reference PyTorch-CUDA eager" figure of SURVEY 8d(i).  fp32 with matmul TF32 off (as the reference leaves it) and on."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

def run(B, tf32, steps=3):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = True      # PyTorch default (only touches convs)
    step, voxels, desc = bench.oracle_step_factory(B, 1, device="cuda")
    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"B": B, "matmul_tf32": tf32, "ms_per_step": ms, "voxels_per_s": voxels / ms * 1e3,
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}

if __name__ == "__main__":
    for tf32 in (False, True):
        print(json.dumps(run(4, tf32)))
