import sys; sys.path.insert(0,'/root/repo')
import torch, ctypes
from audiocodecs_b200 import _lib
L=_lib.lib()
torch.zeros(1, device="cuda")
for cs in (2,4,8,16):
    for smem in (48*1024, 120*1024, 200*1024):
        print("cluster", cs, "smem", smem//1024, "KB ->", L.ac_lstm_tc_max_clusters(cs, smem), _lib.lib().ac_last_error().decode() if L.ac_lstm_tc_max_clusters(cs, smem) < 0 else "")
