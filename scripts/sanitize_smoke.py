import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch, util
cloud, view = util.case("cone_trained_small")
o = util.ours_raster_forward(cloud, view)
dL = np.random.RandomState(5).randn(view.image_height, view.image_width).astype(np.float32)
g = util.ours_raster_backward(cloud, view, o, dL)
from r2_gaussian_b200 import scene
cl = scene.make_cloud(1500, kind="trained", seed=9)
v = util.ours_voxel_forward(cl, (32,32,32),(2.,2.,2.),(0.,0.,0.))
gv = util.ours_voxel_backward(cl, (32,32,32),(2.,2.,2.),(0.,0.,0.), v, np.random.RandomState(1).randn(32,32,32).astype(np.float32))
torch.cuda.synchronize(); print("done", o["R"], v["R"])
