import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from millieye_b200 import configs, radar
from millieye_b200.my_models import Network, define_yolo
from oracle import synth
dev = torch.device("cuda:0")
model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.2).eval()
model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=0, **bench.FUSION_WEIGHTS))
model.to(dev)
inp = bench._fusion_inputs(32, dev, 100)
def step():
    maps = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
    return model(inp["imgs_dev"], maps, inp["boxes"].clone(), 0)
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = step(); torch.cuda.synchronize()
    print(f"call {i}: {(time.perf_counter()-t0)*1e3:.3f} ms rows {out.shape[0]}", flush=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize(); print(f"GRAPHS={os.environ.get('ME_FUSION_GRAPHS','1')}: {(time.perf_counter()-t0)/50*1e3:.3f} ms / call")
