import numpy as np, sys
sys.path.insert(0,'/root/repo')
import front_end_b200 as fe
from front_end_b200 import synth
h,w,P=720,1280,4
Ls,Rs=synth.stereo_batch(h,w,P,seed0=0,n_scenes=2)
f=fe.FrontEnd(max_width=w,max_height=h,max_pairs=P,max_keypoints=8192,n_features=5000,orientation=False,surf_upright=True)
f.set_batch_descriptor(fe.DESC_SURF128)
cb=fe.match_cfg(mode=fe.MATCH_CROSSCHECK,mask=fe.MASK_NONE,norm=fe.NORM_L2)
f.batch_upload(Ls,Rs)
f.batch_run(None,cb,sync=True)
