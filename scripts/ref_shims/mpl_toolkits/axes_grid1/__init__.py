from _absent import Absent as _A

make_axes_locatable = _A("mpl_toolkits.axes_grid1.make_axes_locatable")
