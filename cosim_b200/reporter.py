"""Evaluation report: the reference's `Reporter` interface on top of the batched engine.

Mirrors `core/reporter.py:197-218` -- `Reporter(report_path, config)`, `write_info(info)` once per control step (same
`history` / `timesteps` bookkeeping), `generate_report()` -- and the reference's four sections (`core/reporter.py:252-725`):
set points vs. states, command inputs vs. measured outputs, action oscillation + applied torques + torque distribution,
configuration table.  A fifth section is new: population statistics of a batched run (`BatchedEnv.stats()`, NCCL
all-reduced over ranks), set with `write_population()`.

The reference renders through matplotlib, which is not part of this image; the pages here are drawn by a small PDF writer
(vector line plots, Helvetica text, Flate-compressed content streams), so the report needs nothing beyond numpy.  `info` values
may be python scalars, numpy arrays or torch tensors; for a batched env pass `Info` rows of one traced env
(`Reporter.write_info(env_info, env_index=i)`).
"""
import math
import time
import zlib

import numpy as np

PALETTE = [(0.39, 0.40, 0.95), (0.02, 0.71, 0.83), (0.13, 0.77, 0.37), (0.98, 0.45, 0.09), (0.93, 0.28, 0.60),
           (0.66, 0.33, 0.97), (0.05, 0.65, 0.91), (0.07, 0.09, 0.15)]
INK, MUTED, GRID, BORDER = (0.04, 0.07, 0.13), (0.39, 0.45, 0.55), (0.90, 0.91, 0.92), (0.80, 0.84, 0.88)
DASHES = ["[] 0", "[6 3] 0", "[6 2 1.5 2] 0", "[1.5 2] 0"]
PAGE_W, PAGE_H = 595.0, 842.0      # A4 portrait, points


def _esc(s):
    return str(s).replace("\\", "\\\\").replace("(", "\\(").replace(")", "\\)").encode("latin-1", "replace").decode("latin-1")


class _Page:
    def __init__(self):
        self.ops = []

    def color(self, rgb, stroke=True):
        self.ops.append("%.3f %.3f %.3f %s" % (rgb[0], rgb[1], rgb[2], "RG" if stroke else "rg"))

    def rect(self, x, y, w, h, fill=None, stroke=None, width=0.6):
        if fill is not None:
            self.color(fill, stroke=False)
        if stroke is not None:
            self.color(stroke)
            self.ops.append("%.2f w" % width)
        self.ops.append("%.2f %.2f %.2f %.2f re %s" % (x, y, w, h, "B" if fill is not None and stroke is not None else ("f" if fill is not None else "S")))

    def polyline(self, xs, ys, rgb, width=1.0, dash=0):
        if len(xs) < 2:
            return
        self.color(rgb)
        self.ops.append("%.2f w %s d 1 j" % (width, DASHES[dash % len(DASHES)]))
        pts = ["%.2f %.2f m" % (xs[0], ys[0])] + ["%.2f %.2f l" % (x, y) for x, y in zip(xs[1:], ys[1:])]
        self.ops.append(" ".join(pts) + " S")
        self.ops.append("[] 0 d")

    def text(self, x, y, s, size=9, rgb=INK, bold=False, align="left"):
        s = str(s)
        if align != "left":
            w = 0.5 * size * len(s) * (1.06 if bold else 1.0)
            x = x - w if align == "right" else x - 0.5 * w
        self.color(rgb, stroke=False)
        self.ops.append("BT /%s %.1f Tf %.2f %.2f Td (%s) Tj ET" % ("F2" if bold else "F1", size, x, y, _esc(s)))

    def stream(self):
        return "\n".join(self.ops).encode("latin-1")


class _Pdf:
    """Just enough of PDF 1.4: pages of vector graphics with the two base-14 Helvetica fonts."""

    def __init__(self):
        self.pages = []

    def page(self):
        p = _Page()
        self.pages.append(p)
        return p

    def save(self, path):
        objs = []          # object bodies, 1-based ids in order

        def add(body):
            objs.append(body)
            return len(objs)
        cat, pages_id = add(None), add(None)
        f1 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica /Encoding /WinAnsiEncoding >>")
        f2 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica-Bold /Encoding /WinAnsiEncoding >>")
        kids = []
        for p in self.pages:
            data = zlib.compress(p.stream())
            c = add(b"<< /Length %d /Filter /FlateDecode >>\nstream\n" % len(data) + data + b"\nendstream")
            kids.append(add(("<< /Type /Page /Parent %d 0 R /MediaBox [0 0 %.0f %.0f] /Contents %d 0 R /Resources << /Font << /F1 %d 0 R /F2 %d 0 R >> >> >>"
                             % (pages_id, PAGE_W, PAGE_H, c, f1, f2)).encode()))
        objs[cat - 1] = ("<< /Type /Catalog /Pages %d 0 R >>" % pages_id).encode()
        objs[pages_id - 1] = ("<< /Type /Pages /Count %d /Kids [%s] >>" % (len(kids), " ".join("%d 0 R" % k for k in kids))).encode()
        out = bytearray(b"%PDF-1.4\n%\xe2\xe3\xcf\xd3\n")
        offsets = []
        for i, body in enumerate(objs):
            offsets.append(len(out))
            out += b"%d 0 obj\n" % (i + 1) + body + b"\nendobj\n"
        xref = len(out)
        out += b"xref\n0 %d\n0000000000 65535 f \n" % (len(objs) + 1)
        for o in offsets:
            out += b"%010d 00000 n \n" % o
        out += b"trailer\n<< /Size %d /Root %d 0 R >>\nstartxref\n%d\n%%%%EOF\n" % (len(objs) + 1, cat, xref)
        with open(path, "wb") as f:
            f.write(out)


def _nice_ticks(lo, hi, n=5):
    if not (math.isfinite(lo) and math.isfinite(hi)):
        lo, hi = 0.0, 1.0
    if hi - lo < 1e-12:
        pad = 0.5 if lo == 0 else abs(lo) * 0.1
        lo, hi = lo - pad, hi + pad
    raw = (hi - lo) / n
    mag = 10 ** math.floor(math.log10(raw))
    step = min((1, 2, 2.5, 5, 10), key=lambda s: abs(s * mag - raw)) * mag
    t0 = math.floor(lo / step) * step
    ticks = [t0 + i * step for i in range(int((hi - t0) / step) + 2)]
    return ticks[0], ticks[-1], ticks


def _fmt(v):
    return ("%.3g" % v) if abs(v) >= 1e-3 or v == 0 else ("%.1e" % v)


def _axes(pg, box, title, xlabel, series, legend=True, bars=None):
    """series: [(label, x, y, rgb, width, dash)]; bars: (edges, counts) draws a histogram instead of lines."""
    x0, y0, w, h = box
    pg.text(x0, y0 + h + 6, title, size=10, bold=True)
    if bars is not None:
        edges, counts = bars
        xlo, xhi, ylo, yhi = float(edges[0]), float(edges[-1]), 0.0, float(max(counts.max(), 1))
    else:
        xs = np.concatenate([np.asarray(s[1], float) for s in series]) if series else np.zeros(1)
        ys = np.concatenate([np.asarray(s[2], float) for s in series]) if series else np.zeros(1)
        ys = ys[np.isfinite(ys)] if np.isfinite(ys).any() else np.zeros(1)
        xlo, xhi, ylo, yhi = float(xs.min()), float(xs.max()), float(ys.min()), float(ys.max())
    ylo, yhi, yt = _nice_ticks(ylo, yhi)
    if xhi - xlo < 1e-12:
        xhi = xlo + 1.0
    pg.rect(x0, y0, w, h, fill=(1, 1, 1), stroke=BORDER)
    for t in yt:
        yy = y0 + (t - ylo) / (yhi - ylo) * h
        pg.polyline([x0, x0 + w], [yy, yy], GRID, 0.4)
        pg.text(x0 - 3, yy - 2.5, _fmt(t), size=6.5, rgb=MUTED, align="right")
    _, _, xt = _nice_ticks(xlo, xhi, 6)
    for t in xt:
        if xlo <= t <= xhi:
            xx = x0 + (t - xlo) / (xhi - xlo) * w
            pg.polyline([xx, xx], [y0, y0 + h], GRID, 0.4)
            pg.text(xx, y0 - 9, _fmt(t), size=6.5, rgb=MUTED, align="center")
    pg.text(x0 + w / 2, y0 - 19, xlabel, size=7.5, rgb=MUTED, align="center")
    if bars is not None:
        edges, counts = bars
        for i, c in enumerate(counts):
            bx0 = x0 + (edges[i] - xlo) / (xhi - xlo) * w
            bx1 = x0 + (edges[i + 1] - xlo) / (xhi - xlo) * w
            pg.rect(bx0, y0, max(bx1 - bx0 - 0.3, 0.2), (c - ylo) / (yhi - ylo) * h, fill=PALETTE[0])
        return
    for label, x, y, rgb, width, dash in series:
        x, y = np.asarray(x, float), np.asarray(y, float)
        if len(x) > 1500:                      # decimate long traces: keep min / max of each bucket
            k = int(math.ceil(len(x) / 750))
            n = len(x) // k * k
            yy = y[:n].reshape(-1, k)
            x = np.repeat(x[:n].reshape(-1, k)[:, 0], 2)
            y = np.stack([yy.min(1), yy.max(1)], 1).reshape(-1)
        px = x0 + (x - xlo) / (xhi - xlo) * w
        py = y0 + (np.clip(y, ylo, yhi) - ylo) / (yhi - ylo) * h
        ok = np.isfinite(py)
        pg.polyline(px[ok], py[ok], rgb, width, dash)
    if legend and series:
        ly = y0 + h - 9
        for label, _, _, rgb, width, dash in series[:8]:
            pg.polyline([x0 + 6, x0 + 22], [ly + 2.5, ly + 2.5], rgb, width, dash)
            pg.text(x0 + 26, ly, label, size=6.5, rgb=MUTED)
            ly -= 9


def _header(pg, title, page_no=None):
    pg.rect(0, PAGE_H - 54, PAGE_W, 54, fill=INK)
    pg.rect(0, PAGE_H - 57, PAGE_W, 3, fill=PALETTE[0])
    pg.text(36, PAGE_H - 34, title, size=15, rgb=(1, 1, 1), bold=True)
    if page_no is not None:
        pg.text(PAGE_W - 36, PAGE_H - 33, "%d" % page_no, size=10, rgb=(0.8, 0.84, 0.9), align="right")


def _scalar(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return v


class Reporter:
    def __init__(self, report_path, config):
        self.report_path = report_path
        self.config = config
        self.history = {}
        self.timesteps = 0
        self.population = None

    def write_info(self, info, env_index=None):
        """Append one control step of logged info (core/reporter.py:210-218).  With `env_index`, `info` is the batched env's
        info mapping and row `env_index` of every entry is traced."""
        self.timesteps += 1
        for key in info.keys():
            value = _scalar(info[key])
            if env_index is not None and hasattr(value, "shape") and len(value.shape) >= 1:
                value = value[env_index]
            self.history.setdefault(key, []).append(value)

    def write_population(self, stats):
        """Population statistics of a batched run: the dict `BatchedEnv.stats()` returns (extension)."""
        self.population = {k: float(v) for k, v in stats.items()}

    def _build_config_rows(self, config, indent=0):
        rows = []
        pad = "    " * indent
        for key, value in config.items():
            if isinstance(value, dict):
                rows.append([f"{pad}{key}", ""])
                rows.extend(self._build_config_rows(value, indent + 1))
            elif isinstance(value, (list, tuple)):
                rows.append([f"{pad}{key}", ", ".join(map(str, value))])
            else:
                rows.append([f"{pad}{key}", str(value)])
        return rows

    def _array(self, key):
        return np.array([np.asarray(v, dtype=float) for v in self.history[key]], dtype=float)

    def generate_report(self):
        pdf = _Pdf()
        dt = float(np.asarray(self.history.get("dt", [1])[0]))
        times = np.arange(self.timesteps) * dt
        page_no = [0]

        def new_page(title):
            page_no[0] += 1
            pg = pdf.page()
            _header(pg, title, page_no[0])
            return pg
        # ---- cover
        pg = pdf.page()
        pg.rect(0, 0, PAGE_W, PAGE_H, fill=(0.97, 0.98, 0.99))
        pg.rect(36, PAGE_H - 170, PAGE_W - 72, 110, fill=INK)
        pg.rect(36, PAGE_H - 174, PAGE_W - 72, 4, fill=PALETTE[0])
        pg.text(56, PAGE_H - 110, "Sim-to-Sim Evaluation Report", size=22, rgb=(1, 1, 1), bold=True)
        env_id = str(self.config.get("env", {}).get("id", "")) if isinstance(self.config, dict) else ""
        pg.text(56, PAGE_H - 140, env_id, size=12, rgb=(0.8, 0.84, 0.9))
        y = PAGE_H - 230
        for label, value in (("Generated", time.strftime("%Y-%m-%d %H:%M:%S")), ("Control steps", self.timesteps),
                             ("Duration", "%.2f s" % (self.timesteps * dt)), ("Control period", "%.4f s" % dt)):
            pg.text(56, y, label, size=10, rgb=MUTED)
            pg.text(200, y, value, size=10, bold=True)
            y -= 20
        # ---- 1) set points vs. states
        if "set_points" in self.history and "state" in self.history and self.timesteps:
            sp, st = self._array("set_points"), self._array("state")
            ndim = min(sp.shape[1], st.shape[1])
            per_page, cols = 8, 2
            for start in range(0, ndim, per_page):
                pg = new_page("Set Points vs. States")
                for i, d in enumerate(range(start, min(start + per_page, ndim))):
                    r, c = divmod(i, cols)
                    box = (52 + c * 270, PAGE_H - 230 - r * 178, 225, 120)
                    _axes(pg, box, f"Dimension {d}", "Time (s)", [("Set point", times, sp[:, d], PALETTE[0], 1.1, 0), ("State", times, st[:, d], PALETTE[3], 1.1, 2)])
        # ---- 2) command inputs vs. measured outputs
        cmd_keys = sorted(k for k in self.history if k.startswith("user_command_"))
        measured = [(k, label, unit) for k, label, unit in (("lin_vel_x", "Linear Velocity X", "m/s"), ("lin_vel_y", "Linear Velocity Y", "m/s"),
                                                           ("ang_vel_yaw", "Angular Velocity Yaw", "rad/s")) if k in self.history]
        if measured and cmd_keys:
            pg = new_page("Command Inputs vs. Measured Outputs")
            for i, (k, label, unit) in enumerate(measured):
                series = [(f"Command {ck.replace('user_command_', '')}", times, self._array(ck).reshape(self.timesteps, -1)[:, 0], PALETTE[j % len(PALETTE)], 1.0, j % 4)
                          for j, ck in enumerate(cmd_keys)]
                series.append((f"{label} ({unit})", times, self._array(k).reshape(self.timesteps, -1)[:, 0], (0, 0, 0), 1.8, 0))
                _axes(pg, (52, PAGE_H - 270 - i * 240, 495, 170), f"{label} ({unit})", "Time (s)", series)
        # ---- 3) action oscillation and torques
        if "action_diff_RMSE" in self.history and "torque" in self.history and self.timesteps:
            pg = new_page("Action Oscillation and Applied Torques")
            diffs = self._array("action_diff_RMSE").reshape(-1)
            series = [("da (RMSE)", times, diffs, PALETTE[0], 1.0, 0)]
            win = max(1, min(50, self.timesteps // 10))
            if win > 1:
                ma = np.convolve(diffs, np.ones(win) / win, mode="same")
                series.append((f"da (RMSE) moving average (window={win})", times, ma, PALETTE[3], 1.4, 1))
            _axes(pg, (52, PAGE_H - 270, 495, 170), "Action Oscillation", "Time (s)", series)
            tq = self._array("torque").reshape(self.timesteps, -1)
            _axes(pg, (52, PAGE_H - 510, 495, 170), "Applied Torque of Each Joint", "Time (s)",
                  [(f"Torque {i}", times, tq[:, i], PALETTE[i % len(PALETTE)], 0.9, (i // len(PALETTE)) % 4) for i in range(tq.shape[1])])
            counts, edges = np.histogram(tq.reshape(-1), bins=40)
            _axes(pg, (52, PAGE_H - 750, 495, 170), "Torque Distribution of All Joints", "Torque (Nm)", [], bars=(edges, counts))
        # ---- 4) population statistics (batched runs)
        if self.population:
            pg = new_page("Population Statistics")
            y = PAGE_H - 100
            for k, v in self.population.items():
                pg.text(56, y, k, size=9.5, rgb=MUTED)
                pg.text(320, y, _fmt(v), size=9.5, bold=True)
                pg.polyline([52, PAGE_W - 52], [y - 5, y - 5], GRID, 0.4)
                y -= 18
        # ---- 5) configuration
        rows = self._build_config_rows(self.config) if isinstance(self.config, dict) else []
        per_page = 52
        for start in range(0, len(rows), per_page):
            pg = new_page("Configuration")
            y = PAGE_H - 90
            for name, value in rows[start:start + per_page]:
                indent = (len(name) - len(name.lstrip(" "))) // 4
                pg.text(56 + 12 * indent, y, name.strip(), size=8.5, bold=(value == ""))
                pg.text(300, y, value if len(value) <= 60 else value[:57] + "...", size=8.5, rgb=MUTED)
                pg.polyline([52, PAGE_W - 52], [y - 4, y - 4], GRID, 0.3)
                y -= 14
        pdf.save(self.report_path)
        return self.report_path
