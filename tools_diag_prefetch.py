import sys, threading, faulthandler, os
faulthandler.enable()
sys.path.insert(0, '.')
sys.argv = ['x', '--stacks', '1', '--batch', '32', '--steps', '3']
import tools_pipeline_bench as t
import hgb200
orig_close = hgb200.dataset_builder.Prefetcher.close
def close(self):
    orig_close(self)
    print("after close: alive =", self._thread.is_alive(), [th.name for th in threading.enumerate()], flush=True)
hgb200.dataset_builder.Prefetcher.close = close
t.main()
print("main done; threads:", [(th.name, th.daemon) for th in threading.enumerate()], flush=True)
import torch
torch.cuda.synchronize()
print("exiting", flush=True)
