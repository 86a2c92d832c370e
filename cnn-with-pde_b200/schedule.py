"""Host-side logic shared by the layers: the per-sweep time schedule of the implicit family.

The reference accumulates ``current_time`` in Python double by repeated ``+= dt / 2`` and hands
``t``, ``dt / 2`` (or ``dt``) and ``dx ** 2`` to ATen as Python scalars, which rounds them to the
tensor dtype (fp32).  The kernels take the same fp32 values, so the schedule is rebuilt here
with the same accumulation (mnist_test.py:49-63, cifar10.py:84-110, cifar_2version.py:81-101).
"""
from __future__ import annotations

import numpy as np

from . import _cabi


def adi_schedule_lists(steps: int, dt: float, hx: float, hy: float, lie: bool):
    """Python-double schedule: (t, dts, h) per sweep, sweeps ordered as executed."""
    t_list, dts_list, h_list = [], [], []
    current_time = 0.0
    for _ in range(steps):
        if not lie:
            t_list.append(current_time); dts_list.append(dt / 2); h_list.append(hx)
            current_time += dt / 2
            t_list.append(current_time); dts_list.append(dt); h_list.append(hy)
            current_time += dt / 2
            t_list.append(current_time); dts_list.append(dt / 2); h_list.append(hx)
        else:
            t_list.append(current_time); dts_list.append(dt / 2); h_list.append(hx)
            current_time += dt / 2
            t_list.append(current_time); dts_list.append(dt / 2); h_list.append(hy)
            current_time += dt / 2
    return t_list, dts_list, h_list


def adi_schedule(steps: int, dt: float, hx: float, hy: float, lie: bool) -> "_cabi.AdiSchedule":
    t_list, dts_list, h_list = adi_schedule_lists(steps, dt, hx, hy, lie)
    if len(t_list) > _cabi.MAX_SWEEPS:
        raise ValueError(f"num_steps={steps} needs {len(t_list)} sweeps; this build supports {_cabi.MAX_SWEEPS}")
    s = _cabi.AdiSchedule()
    for i, (t, d, h) in enumerate(zip(t_list, dts_list, h_list)):
        s.t[i] = float(np.float32(t))
        s.dts[i] = float(np.float32(d))
        s.h2[i] = float(np.float32(h ** 2))
    return s
