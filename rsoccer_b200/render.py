"""Picture of ONE match of a batch as an RGB array (host side, numpy only).

Stands in for rsoccer_gym/Render (pygame surfaces drawn by `VSSBaseEnv.render` /
`SSLBaseEnv.render`, vss_gym_base.py:148-187) for `render_mode="rgb_array"`: same picture
elements -- green background, field outline with the goals, centre line and circle, penalty
areas, orange ball, blue / yellow robots with a heading mark (squares for VSS, discs with a flat
kicker side for SSL) -- and the reference's palette (Render/utils.py:2-15).  No window, no
pygame: `render_mode="human"` is not offered.  The input is one row of the wire state
(`BatchedWorld.get_state()[i]`, Entities/Frame.py layout: m, degrees, field-centred), so it runs
without a GPU.
"""
import math

import numpy as np

BG_GREEN = (20, 90, 45)
WHITE = (220, 220, 220)
ORANGE = (253, 106, 2)
BLUE = (0, 64, 255)
YELLOW = (250, 218, 94)
ROBOT_BLACK = (25, 25, 25)


class _Canvas:
    def __init__(self, field, width_px, margin_m):
        self.L, self.W = float(field["length"]), float(field["width"])
        self.gd = float(field["goal_depth"])
        self.margin = margin_m + self.gd
        self.scale = width_px / (self.L + 2 * self.margin)
        self.w = int(width_px)
        self.h = int(round((self.W + 2 * self.margin) * self.scale))
        self.img = np.empty((self.h, self.w, 3), np.uint8)
        self.img[:] = BG_GREEN
        ys, xs = np.mgrid[0:self.h, 0:self.w]
        # pixel centres in field coordinates (x right, y up)
        self.X = (xs + 0.5) / self.scale - self.L / 2 - self.margin
        self.Y = self.W / 2 + self.margin - (ys + 0.5) / self.scale
        self.px = 1.0 / self.scale

    def fill(self, mask, colour):
        self.img[mask] = colour

    def rect_outline(self, x0, y0, x1, y1, colour=WHITE):
        t = 0.75 * self.px
        inside = (self.X >= x0 - t) & (self.X <= x1 + t) & (self.Y >= y0 - t) & (self.Y <= y1 + t)
        core = (self.X > x0 + t) & (self.X < x1 - t) & (self.Y > y0 + t) & (self.Y < y1 - t)
        self.fill(inside & ~core, colour)

    def circle_outline(self, cx, cy, r, colour=WHITE):
        d = np.hypot(self.X - cx, self.Y - cy)
        self.fill(np.abs(d - r) <= 0.75 * self.px, colour)

    def disc(self, cx, cy, r, colour):
        self.fill((self.X - cx) ** 2 + (self.Y - cy) ** 2 <= r * r, colour)

    def local(self, cx, cy, theta_deg):
        c, s = math.cos(math.radians(theta_deg)), math.sin(math.radians(theta_deg))
        dx, dy = self.X - cx, self.Y - cy
        return c * dx + s * dy, -s * dx + c * dy          # robot frame: x forward, y left


def render_rgb(state_row, field, kind, n_blue, n_yellow, width_px=750, margin_m=0.1):
    """state_row: 5 + K (n_blue + n_yellow) floats of one match (K = 6 VSS, 11 SSL);
    field: dict with the reference's Field keys (Entities/Field.py:4-21); kind: "vss" | "ssl".
    Returns uint8 [H, width_px, 3]."""
    st = np.asarray(state_row, dtype=np.float64).reshape(-1)
    K = 6 if kind == "vss" else 11
    R = n_blue + n_yellow
    if st.size != 5 + K * R:
        raise ValueError("state row of %d floats, expected %d" % (st.size, 5 + K * R))
    cv = _Canvas(field, width_px, margin_m)
    L2, W2 = cv.L / 2, cv.W / 2
    gw2, gd = float(field["goal_width"]) / 2, cv.gd
    pl, pw2 = float(field["penalty_length"]), float(field["penalty_width"]) / 2
    # field lines
    cv.rect_outline(-L2, -W2, L2, W2)
    cv.rect_outline(-L2 - gd, -gw2, -L2, gw2)
    cv.rect_outline(L2, -gw2, L2 + gd, gw2)
    cv.rect_outline(-L2, -pw2, -L2 + pl, pw2)
    cv.rect_outline(L2 - pl, -pw2, L2, pw2)
    cv.fill((np.abs(cv.X) <= 0.75 * cv.px) & (np.abs(cv.Y) <= W2), WHITE)
    cv.circle_outline(0.0, 0.0, 0.2 if kind == "vss" else 0.5)
    # robots
    rr = float(field["rbt_radius"])
    for k in range(R):
        x, y, th = st[5 + K * k: 8 + K * k]
        team = BLUE if k < n_blue else YELLOW
        lx, ly = cv.local(x, y, th)
        if kind == "vss":
            body = (np.abs(lx) <= rr) & (np.abs(ly) <= rr)                 # 7.5 cm cube seen from above
            tag = body & (lx >= 0.2 * rr)
        else:
            dk = float(field["rbt_distance_center_kicker"])
            body = (lx * lx + ly * ly <= rr * rr) & (lx <= dk)            # disc with the flat kicker side
            tag = body & (lx * lx + ly * ly <= (0.55 * rr) ** 2)
        cv.fill(body, ROBOT_BLACK)
        cv.fill(tag, team)
        cv.fill(body & (np.abs(ly) <= 0.75 * cv.px) & (lx >= 0), WHITE)    # heading mark
    # ball last (on top, as in the reference's draw order)
    cv.disc(st[0], st[1], float(field["ball_radius"]), ORANGE)
    return cv.img
