import torch, sys
sys.path.insert(0, ".")
from lanpaint_b200 import hostnoise
dev = torch.device("cuda:0")
hostnoise._draw(128*4*128*128, 1, dev); torch.cuda.synchronize()
hostnoise._draw(128*4*128*128, 2, dev); torch.cuda.synchronize()
