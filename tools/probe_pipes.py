"""Instruction-rate probes of the non-tensor pipes (lkf_probe_pipe): warp instructions per clock per SM for FFMA, IMAD,
IDP.2A, SHF+LOP3 and FFMA+SHF, at the SM clock nvidia-smi reports under load.  Usage: probe_pipes.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200.device import Workspace, stream_ptr

ws = Workspace.current()
out = torch.zeros(4, device="cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
blocks, it = sms * 16, 2000
names = {0: "FFMA", 1: "IMAD", 2: "IDP.2A", 3: "SHF+LOP3", 4: "FFMA+SHF", 5: "IMAD/SHF", 6: "IMAD/FFMA", 7: "IDP/SHF/FFMA"}
for kind in range(8):
    for _ in range(2):
        ws.lib.lkf_probe_pipe(kind, blocks, it, out.data_ptr(), stream_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ws.lib.lkf_probe_pipe(kind, blocks, it, out.data_ptr(), stream_ptr())
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True,
                               text=True).stdout.split()[0])
    per_thread = it * 128 * (2 if kind in (3, 4) else 1)
    warp_inst = blocks * 8 * per_thread
    rate = warp_inst / (ms * 1e-3) / (mhz * 1e6) / sms
    print(f"{names[kind]:9s} {ms:8.3f} ms  {rate:5.2f} warp-instructions/clk/SM ({rate * 32:6.1f} lanes/clk/SM) at {mhz:.0f} MHz"
          + (f"  = {blocks * 256 * per_thread * 2 / (ms * 1e-3) / 1e12:6.2f} TFLOP/s" if kind == 0 else ""))
