import torch, sys
sys.path.insert(0,'/root/repo')
import pydrodem_b200 as hc
from pydrodem_b200 import synth, _lib
n=int(sys.argv[1]) if len(sys.argv)>1 else 10801
dev=torch.device('cuda')
z=torch.empty((n,n),dtype=torch.float32,device=dev)
for r0 in range(0,n,1024):
    nr=min(1024,n-r0); z[r0:r0+nr]=synth.make_dem(n,n,1002,row0=r0,nrows=nr,device=dev)
out=hc.fill_d8_accum(z,dx=30.,dy=30.,seed_mode=1,table=1,want_slope=False)
_lib.debug_read()
out=hc.fill_d8_accum(z,dx=30.,dy=30.,seed_mode=1,table=1,want_slope=False)
d=_lib.debug_read()
tiles=((n+63)//64)**2
print('minedge tiles visited per round', d[6:11], 'dense slow visits', d[5]); print('tiles',tiles,'fast sweeps',d[0],'fast tiles w/ flats',d[1],'avg sweeps',d[0]/max(d[1],1),'pending tiles',d[2],'slow sweeps',d[3],'slow visits',d[4], 'avg', d[3]/max(d[4],1))

nb=max(d[15],1)
print('acc_tile_local cycles/tile: load',d[11]//nb,'classify+scatter',d[12]//nb,'walk',d[13]//nb,'epilogue',d[14]//nb)
