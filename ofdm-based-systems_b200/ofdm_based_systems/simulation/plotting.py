"""Constellation image for ``results["constellation_plot"]`` drawn with Pillow.

The reference renders this with matplotlib (simulation/models.py:622-799); matplotlib is not part of
this image, and a 1e9-bit run cannot scatter every symbol anyway, so the picture is drawn directly:
received symbols as translucent blue dots, the ideal points in red, the BER / SNR / PAPR box, and for
adaptive runs a second panel with the constellation-order histogram."""
from __future__ import annotations

from typing import Optional

import numpy as np
from PIL import Image, ImageDraw

_SIZE = 800
_MARGIN = 70
_LIMIT = 1.5
_MAX_POINTS = 200_000


def _to_px(values: np.ndarray, flip: bool = False) -> np.ndarray:
    span = _SIZE - 2 * _MARGIN
    unit = (np.clip(values, -_LIMIT, _LIMIT) + _LIMIT) / (2 * _LIMIT)
    if flip:
        unit = 1.0 - unit
    return _MARGIN + unit * span


def _scatter_panel(received: np.ndarray, ideal: np.ndarray, title: str, text: str) -> Image.Image:
    img = Image.new("RGB", (_SIZE, _SIZE), "white")
    # density accumulation gives the alpha = 0.1 look without per-point compositing
    pts = np.asarray(received).reshape(-1)
    pts = pts[np.isfinite(pts.real) & np.isfinite(pts.imag)]
    if pts.size > _MAX_POINTS:
        pts = pts[:: pts.size // _MAX_POINTS + 1]
    if pts.size:
        x = _to_px(pts.real).astype(np.int64)
        y = _to_px(pts.imag, flip=True).astype(np.int64)
        hist = np.zeros((_SIZE, _SIZE), dtype=np.float64)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                np.add.at(hist, (np.clip(y + dy, 0, _SIZE - 1), np.clip(x + dx, 0, _SIZE - 1)), 1.0)
        alpha = 1.0 - 0.9 ** hist                       # n overlapping dots of alpha 0.1
        rgb = np.full((_SIZE, _SIZE, 3), 255.0)
        blue = np.array([0.0, 0.0, 255.0])
        rgb = rgb * (1 - alpha[..., None]) + blue * alpha[..., None]
        img = Image.fromarray(rgb.astype(np.uint8), "RGB")
    draw = ImageDraw.Draw(img)
    lo, hi = _MARGIN, _SIZE - _MARGIN
    for g in np.linspace(-_LIMIT, _LIMIT, 7):
        gx = float(_to_px(np.array([g]))[0])
        draw.line([(gx, lo), (gx, hi)], fill=(220, 220, 220))
        draw.line([(lo, gx), (hi, gx)], fill=(220, 220, 220))
        draw.text((gx - 10, hi + 6), f"{g:.1f}", fill="black")
        draw.text((lo - 34, _SIZE - gx - 6), f"{g:.1f}", fill="black")
    mid = float(_to_px(np.array([0.0]))[0])
    draw.line([(mid, lo), (mid, hi)], fill="black")
    draw.line([(lo, mid), (hi, mid)], fill="black")
    draw.rectangle([lo, lo, hi, hi], outline="black")
    for p in np.asarray(ideal).reshape(-1):
        px, py = float(_to_px(np.array([p.real]))[0]), float(_to_px(np.array([p.imag]), flip=True)[0])
        draw.ellipse([px - 4, py - 4, px + 4, py + 4], fill=(255, 0, 0))
    draw.text((_SIZE // 2 - 3 * len(title), 20), title, fill="black")
    draw.text((_SIZE // 2 - 24, _SIZE - 28), "In-Phase", fill="black")
    draw.text((8, _SIZE // 2), "Q", fill="black")
    lines = text.split("\n")
    draw.rectangle([lo + 10, lo + 10, lo + 150, lo + 18 + 14 * len(lines)], fill="white", outline="gray")
    for i, line in enumerate(lines):
        draw.text((lo + 16, lo + 14 + 14 * i), line, fill="black")
    return img


def _order_histogram_panel(orders: np.ndarray, num_subcarriers: int) -> Image.Image:
    img = Image.new("RGB", (_SIZE, _SIZE), "white")
    draw = ImageDraw.Draw(img)
    active = orders[orders > 0]
    values, counts = np.unique(active, return_counts=True) if active.size else (np.array([]), np.array([]))
    lo, hi = _MARGIN, _SIZE - _MARGIN
    draw.rectangle([lo, lo, hi, hi], outline="black")
    draw.text((_SIZE // 2 - 110, 20), "Constellation Order Distribution", fill="black")
    if counts.size:
        width = (hi - lo) / (len(values) * 1.5 + 0.5)
        top = counts.max()
        for i, (v, c) in enumerate(zip(values, counts)):
            x0 = lo + width * (0.5 + 1.5 * i)
            h = (hi - lo - 30) * c / top
            shade = int(60 + 150 * i / max(len(values) - 1, 1))
            draw.rectangle([x0, hi - h, x0 + width, hi], fill=(68, shade, 140), outline="black")
            draw.text((x0 + width / 2 - 8, hi - h - 14), str(int(c)), fill="black")
            draw.text((x0 + width / 2 - 8, hi + 6), str(int(v)), fill="black")
    stats = (f"Total Subcarriers: {num_subcarriers}\nActive: {int(np.sum(orders > 0))}\n"
             f"Inactive: {int(np.sum(orders == 0))}\nAvg Order: {float(active.mean()) if active.size else 0:.1f}")
    for i, line in enumerate(stats.split("\n")):
        draw.text((hi - 170, lo + 10 + 14 * i), line, fill="black")
    return img


def draw_constellation_image(received: np.ndarray, ideal: np.ndarray, *, title: str, ber: float, snr_db: float,
                             papr_db: float, orders: Optional[np.ndarray] = None,
                             num_subcarriers: int = 0) -> Image.Image:
    text = f"BER: {ber:.6f}\nSNR: {snr_db} dB\nPAPR: {papr_db:.2f} dB"
    if orders is None:
        return _scatter_panel(received, ideal, title, text)
    left = _scatter_panel(received, ideal, "Constellation Diagram (Adaptive Modulation)", text)
    right = _order_histogram_panel(np.asarray(orders), num_subcarriers)
    both = Image.new("RGB", (2 * _SIZE, _SIZE), "white")
    both.paste(left, (0, 0))
    both.paste(right, (_SIZE, 0))
    return both
