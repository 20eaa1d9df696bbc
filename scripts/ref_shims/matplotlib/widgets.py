from _absent import Absent as _A

Slider = _A("matplotlib.widgets.Slider")
