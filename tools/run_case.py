import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch
import parity_cases as PC
name = sys.argv[1]
dt = torch.bfloat16 if len(sys.argv) < 3 or sys.argv[2] == 'bf16' else torch.float32
c = PC.ALL_CASES[name]()
fails, worst = c.check(dt, torch.device('cuda:0'), verbose=True)
print("fails", len(fails), "worst", worst)
