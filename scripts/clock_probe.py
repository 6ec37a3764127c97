"""SM clock actually delivered: torch.cuda._sleep(n) spins n clock64 cycles on one SM; CUDA events give the time.
    python scripts/clock_probe.py      idle, after a short burst, and right after ~1 s of the attention kernel back to back"""
import sys, os, math, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unav_yolyolva_b200 import kernels as K

dev = torch.device("cuda", 0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def mhz(n=20_000_000):
    a.record(); torch.cuda._sleep(n); b.record(); torch.cuda.synchronize()
    return n / (a.elapsed_time(b) * 1e3)


def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()


torch.cuda._sleep(1000); torch.cuda.synchronize()
print("idle start:", [round(mhz()) for _ in range(3)], smi())
nb, T, nh, hs = 32, 224, 4, 64
op = K.BF16X2
C = nh * hs
qo = K.new_operand(nb * T, C, op, dev); qo.normal_()
ko = K.new_operand(nb * T, C, op, dev); ko.normal_()
vt = K.new_operand(nb * C, T, op, dev); vt.normal_()
kmask = torch.ones(nb, T, dtype=torch.uint8, device=dev)
out = K.new_operand(nb * T, C, op, dev)
grp = [{"q": qo, "k": ko, "vt": vt, "kmask": kmask, "out": out}]
K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50):
        K.attention_tc(grp, nb, T, T, nh, hs, 1 / math.sqrt(hs), op)
for rep in range(3):
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    print(f"50 attention launches: {a.elapsed_time(b) / 50 * 1e3:.1f} us each; clock right after: {round(mhz(2_000_000))} MHz", smi())
for _ in range(40):
    gr.replay()
torch.cuda.synchronize()
print("after ~50 ms of load:", round(mhz(2_000_000)), smi())
a.record()
for _ in range(800):
    gr.replay()
b.record(); torch.cuda.synchronize()
print(f"800 replays: {a.elapsed_time(b) / 800 / 50 * 1e3:.1f} us per launch; clock right after: {round(mhz(2_000_000))} MHz", smi())
