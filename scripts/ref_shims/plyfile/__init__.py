from _absent import Absent as _A

PlyData = _A("plyfile.PlyData")
PlyElement = _A("plyfile.PlyElement")
