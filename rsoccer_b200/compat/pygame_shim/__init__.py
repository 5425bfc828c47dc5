"""Import-only stand-in for `pygame` (absent from this image).  The reference imports it at
module scope (vss_gym_base.py:12, Render/*.py) but touches it only when render_mode is set;
rendering is out of scope, so any use raises."""


class _Unavailable:
    def __getattr__(self, name):
        raise RuntimeError("pygame is not installed: rendering is out of scope of rsoccer_b200")


SRCALPHA = 0
draw = display = time = event = surfarray = transform = font = image = _Unavailable()


class Surface:
    def __init__(self, *a, **k):
        raise RuntimeError("pygame is not installed: rendering is out of scope of rsoccer_b200")


def init():
    raise RuntimeError("pygame is not installed: rendering is out of scope of rsoccer_b200")


def quit():
    pass
