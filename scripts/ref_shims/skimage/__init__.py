from _absent import Absent as _A

measure = _A("skimage.measure")
