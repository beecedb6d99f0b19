"""One depth frame (640 x 480 PointXYZRGB) through fx_cloud_filter a few times -- ncu launch-list target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
dev = torch.device("cuda:0")
N = 640 * 480
c = torch.zeros((N, 8), dtype=torch.float32, device=dev)
c[:, 0].uniform_(-3.0, 3.0); c[:, 1].uniform_(-2.0, 2.0); c[:, 2].uniform_(-0.5, 5.0)
c[: N // 2, 2] = 3.0 + 0.03 * torch.randn(N // 2, device=dev)
for _ in range(3):
    out, counts = fx.cloud.cloud_filter(c, rgb_offset=4)
torch.cuda.synchronize()
print(counts.tolist())
