import os, sys, time
sys.path.insert(0, '/root/repo')
import threading
import torch, bench, triad_b200
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
from triad_b200 import regularizers as R
cfg = bench.CONFIGS["cfg2"]; dev = torch.device("cuda", 0)
sets = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 3)
for s in sets: s[0].requires_grad_(True); s[1].requires_grad_(True)
m = triad_b200.TriadHotPath(1.5).to(dev)
def step(q, v, mask):
    q.grad = v.grad = m.temperature.grad = None
    clip, tok = m.compute_all_similarities_av(q, v)
    total = m.compute_contrastive_loss_av(clip, tok)[0]
    total.backward()
def run(tag, n):
    rows, stop = [], [False]
    def poll():
        while not stop[0]:
            rows.append((pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM), round(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3),
                         hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(H)), pynvml.nvmlDeviceGetTemperature(H, 0)))
            time.sleep(0.01)
    th = threading.Thread(target=poll, daemon=True); th.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i in range(n):
        step(*sets[i % 3]); evs[i + 1].record()
    torch.cuda.synchronize()
    per = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(n)]
    stop[0] = True; th.join()
    print('   nvml:', rows[::3])
    print(tag, per, 'free GB', torch.cuda.mem_get_info(dev)[0] / 2**30, 'reserved', torch.cuda.memory_reserved() / 2**30, flush=True)
m.triad_regularizers = False
run('contrastive', 25)
time.sleep(2)
run('contrastive burst', 60)
m.triad_regularizers = True
print('merged ok', R.merged_forward_ok(sets[0][0], sets[0][1]))
run('full warm', 2)
run('full', 10)
run('full', 10)
time.sleep(2)
run('full after idle', 40)
