import os, sys
mode = sys.argv[1]
if mode == "inproc":
    os.environ["NCCL_DEBUG"] = "INFO"; os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"; os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import torch, torch.distributed as dist
r = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
x = torch.ones(16, dtype=torch.float64, device="cuda"); dist.all_reduce(x); torch.cuda.synchronize()
if r == 0: print("sum", x[0].item(), flush=True)
dist.destroy_process_group()
