"""Restatement of the few gx drawing routines the reference's example programs use (test infrastructure only).

Why a renderer in a repository whose scope excludes rendering: the reference keeps the PNGs its Go binary produced for
`examples/density` under doc/ (README.md:201-206).  Those images are the only recorded OUTPUT of the reference for the hot
path - kNN (open and periodic) and Density2D with all three kernels on BASELINE configs[0] - so drawing our results the way
the example does and comparing pixels pins the path against the real Go program (tests/test_reference_images.py).

    NewCanvas / Clear / DrawPoint      gx/graphics.go:50-72
    DrawDisk                           gx/graphics.go:93-110   (float32 arithmetic, `<=` on the squared radius)
    DrawCircle                         gx/graphics.go:74-91
    DrawLine / DrawRect                gx/graphics.go:238-288  (only axis-parallel lines are needed here)
    RainbowRamp / HeatRamp / ToxicRamp / ParaRamp    gx/graphics.go:141-236
    MakeTreePlot / PlotCells / PlotBoundingCircles   sim/visualization.go:9-75 (over the oracle's tree dump)
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

f32 = np.float32

# https://www.kennethmoreland.com/color-advice/ tables as listed in gx/graphics.go:168-186 (black body), 201-219 (Kindlmann)
BLACKBODY = [0, 0, 0, 36, 15, 9, 62, 22, 17, 90, 27, 22, 119, 30, 26, 150, 33, 30, 180, 38, 34, 197, 65, 28, 214, 88, 19, 228, 112, 7,
             231, 141, 18, 233, 169, 29, 233, 195, 39, 231, 222, 50, 246, 240, 144, 255, 255, 255, 255, 255, 255]
KINDLMANN = [0, 0, 0, 37, 3, 57, 37, 5, 109, 24, 8, 163, 8, 51, 160, 6, 83, 127, 5, 105, 105, 6, 127, 83, 7, 148, 47, 15, 168, 8,
             63, 186, 9, 133, 199, 10, 205, 205, 10, 251, 210, 163, 253, 232, 223, 255, 255, 255, 255, 255, 255]


def _ramp17(table, index: int):
    t = f32(index % 16) / f32(16)
    i = index // 16
    return tuple(int(f32(table[i * 3 + 3 + c]) * t + f32(table[i * 3 + c]) * (f32(1) - t)) for c in range(3))


def HeatRamp(index: int):
    return _ramp17(BLACKBODY, index)


def ToxicRamp(index: int):
    return _ramp17(KINDLMANN, index)


def ParaRamp(index: int):
    x = float(index)
    r = min(215.0, max(0.0, -abs((x - 60) * 31 / 17 - 232) + 245))
    g = min(190.0, max(0.0, -abs(x * 31 / 17 - 232) + 245))
    b = min(215.0, max(0.0, -abs((x + 70) * 31 / 17 - 232) + 245))
    return int(r), int(g), int(b)


def RainbowRamp(index: int):
    x = int(index)
    r = max(max(min(255, 620 - 4 * x), 0), 2 * x - 400)
    g = max(min(min(255, 3 * x), 820 - 4 * x), 0)
    b = min(max(0, 4 * x - 620), 255)
    return r & 0xFF, g & 0xFF, b & 0xFF


def colour_index(rho: float, scale: float) -> int:
    """uint8(math.Min(float64(rho / scale * 255), 255)), density.go:23-24, 148-149"""
    return int(min(float(rho / scale * 255), 255.0)) & 0xFF


class Canvas:
    """gx.Canvas over an RGB array (every colour the examples draw is opaque; Go's encoder then writes 8-bit RGB)"""

    def __init__(self, width: int, height: int):
        self.W, self.H = width, height
        self.img = np.zeros((height, width, 3), dtype=np.uint8)  # Clear(gx.BLACK)

    def DrawPoint(self, x: int, y: int, colour):
        if 0 <= x < self.W and 0 <= y < self.H:  # image.NRGBA.Set ignores points outside the rectangle
            self.img[y, x] = colour

    def DrawDisk(self, cx, cy, radius, colour):
        cx, cy, radius = f32(cx), f32(cy), f32(radius)
        xa, xb = int(math.floor(float(cx - radius))), int(math.ceil(float(cx + radius)))
        ya, yb = int(math.floor(float(cy - radius))), int(math.ceil(float(cy + radius)))
        rr = radius * radius
        for x in range(xa, xb + 1):
            dx = f32(x) - cx
            for y in range(ya, yb + 1):
                dy = f32(y) - cy
                if dx * dx + dy * dy <= rr:
                    self.DrawPoint(x, y, colour)

    def DrawCircle(self, cx, cy, radius, border, colour, box_with_border=False):
        """box_with_border: the bounding box of the scan covers radius + border.  The current gx.DrawCircle scans
        [c - radius, c + radius] only and so clips the ring at its four extreme points; the pictures under doc/ were drawn
        by a version that did not clip (tests/test_reference_images.py)"""
        cx, cy, radius, border = f32(cx), f32(cy), f32(radius), f32(border)
        reach = radius + border if box_with_border else radius
        xa, xb = int(math.floor(float(cx - reach))), int(math.ceil(float(cx + reach)))
        ya, yb = int(math.floor(float(cy - reach))), int(math.ceil(float(cy + reach)))
        lo, hi = radius * radius, (radius + border) * (radius + border)
        for x in range(max(xa, 0), min(xb, self.W - 1) + 1):
            dx = f32(x) - cx
            for y in range(max(ya, 0), min(yb, self.H - 1) + 1):
                dy = f32(y) - cy
                r2 = dx * dx + dy * dy
                if lo <= r2 <= hi:
                    self.img[y, x] = colour

    def DrawRect(self, lower_left, upper_right, colour):
        (x1, y1), (x2, y2) = lower_left, upper_right
        self.DrawLine((x1, y1), (x2, y1), colour)
        self.DrawLine((x2, y1), (x2, y2), colour)
        self.DrawLine((x2, y2), (x1, y2), colour)
        self.DrawLine((x1, y2), (x1, y1), colour)

    def DrawLine(self, start, end, colour):
        (x0, y0), (x1, y1) = start, end
        if x0 == x1:
            for y in range(min(y0, y1), max(y0, y1) + 1):
                self.DrawPoint(x0, y, colour)
        elif y0 == y1:
            for x in range(min(x0, x1), max(x0, x1) + 1):
                self.DrawPoint(x, y0, colour)
        else:
            raise NotImplementedError("only the axis-parallel separator lines of density.go:90-92 are restated")

    def sha256(self) -> str:
        return hashlib.sha256(self.img.tobytes()).hexdigest()


WHITE = (255, 255, 255)


def draw_density_panel(canvas: Canvas, pos, rho, side: int, position_index: int, ramp):
    """calcAndDrawDensity's drawing loop (density.go:17-38) for particles in the order given (= Root.Particles order)"""
    off_x, off_y = (side * position_index) % canvas.W, (side * position_index) // canvas.W * side
    for p, r in zip(pos, rho):
        x = f32(p[0]) * f32(side) + f32(off_x)
        y = f32(p[1]) * f32(side) + f32(off_y)
        canvas.DrawDisk(x, y, 4, ramp(colour_index(r, 6000)))


def draw_density_test(canvas: Canvas, pos, rho):
    """periodicVisualTest's drawing loop (density.go:143-160): ToxicRamp of rho / 32000, disks of radius 2"""
    for p, r in zip(pos, rho):
        x = f32(p[0]) * f32(canvas.W)
        y = f32(p[1]) * f32(canvas.H)
        canvas.DrawDisk(x, y, 2, ToxicRamp(colour_index(r, 32000)))


ORANGE, GREEN = (255, 165, 0), (0, 255, 0)


class Tree:
    """the oracle's tree dump (oracle.Oracle.tree) with the two recursions of sim/visualization.go"""

    def __init__(self, geo, link):
        self.geo, self.link = geo, link
        self.depth = np.zeros(len(geo), dtype=np.int64)
        for i in range(len(geo) - 1, -1, -1):  # pre-order: children have larger indices.  Depth(), core.go:327-336
            lo, up = link[i, 0], link[i, 1]
            self.depth[i] = max(self.depth[up] if up >= 0 else 0, self.depth[lo] if lo >= 0 else 0) + 1

    def PlotCells(self, canvas: Canvas, node: int, colour_index: int, max_colour_index: int):
        g = self.geo[node]
        x1, y1 = int(g[0] * float(canvas.W)), int(g[1] * float(canvas.H))
        x2, y2 = int(g[2] * float(canvas.W)), int(g[3] * float(canvas.H))
        lo, up = self.link[node, 0], self.link[node, 1]
        if lo >= 0:
            self.PlotCells(canvas, lo, colour_index + 1, max_colour_index)
        if up >= 0:
            self.PlotCells(canvas, up, colour_index + 1, max_colour_index)
        canvas.DrawRect((x1, y1), (x2 - 1, y2 - 1), RainbowRamp((colour_index * 256 // max_colour_index) & 0xFF))

    def PlotBoundingCircles(self, canvas: Canvas, node: int, max_depth: int, colour, box_with_border=False):
        if self.depth[node] <= max_depth:
            g = self.geo[node]
            x, y = f32(g[4] * float(canvas.W)), f32(g[5] * float(canvas.H))
            r = f32(g[6]) * f32(canvas.W)
            if r < 2:
                r = f32(2)
            canvas.DrawCircle(x, y, r, 1.0, colour, box_with_border)
        lo, up = self.link[node, 0], self.link[node, 1]
        if up >= 0:
            self.PlotBoundingCircles(canvas, up, max_depth, colour, box_with_border)
        if lo >= 0:
            self.PlotBoundingCircles(canvas, lo, max_depth, colour, box_with_border)


def MakeTreePlot(tree: Tree, pos, w: int, h: int) -> Canvas:
    """visualization.go:9-30: cells coloured by level, particles as white points, a rainbow strip in the lower left"""
    canvas = Canvas(w, h)
    tree.PlotCells(canvas, 0, 0, int(tree.depth[0]))
    for p in pos:
        canvas.DrawPoint(int(p[0] * float(w)), int(p[1] * float(h)), WHITE)
    for i in range(256):
        canvas.DrawLine((i, h), (i, h - 10), RainbowRamp(i))
    return canvas
