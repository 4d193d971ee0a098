import sys, numpy as np
sys.path.insert(0,'.')
import rustcv_b200 as R
from oracle import pyoracle as O
R.imgproc.init(0)
al=int(sys.argv[1]) if len(sys.argv)>1 else 4
R.imgproc.set_option("warp.box_align", al)
R.imgproc.set_option("warp.x_align", int(sys.argv[2]) if len(sys.argv)>2 else 4)
h,w=128,128
a=O.fill_f32(70,h*w).reshape(h,w)
M=R.imgproc.get_rotation_matrix_2d(((w-1)/2,(h-1)/2),15.0,1.0)
s=R.Mat.from_numpy(a).upload(); d=s.like()
R.imgproc.warp_affine(s,d,M,border_value=0.25)
got=d.to_numpy(); want=O.warp_affine(a,M.ravel(),border_value=0.25)
print("align",al,"max abs diff",np.abs(got-want).max())
