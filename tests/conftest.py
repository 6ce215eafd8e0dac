import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """code-index parity of every oracle / golden comparison that ran (tests/common.py:log_parity)"""
    try:
        from common import PARITY_LOG
    except Exception:
        return
    if not PARITY_LOG:
        return
    tr = terminalreporter
    tr.write_sep("-", "code-index parity vs oracle / reference goldens")
    total = flips = 0
    for r in PARITY_LOG:
        total += r["total"]
        flips += r["flips"]
        mm = "" if r["min_margin"] is None else f"  (oracle min top-2 margin {r['min_margin']:.2e})"
        extra = f"  margins at flips {r['flip_margins']}" if r["flips"] else ""
        tr.write_line(f"{r['case']}: {r['flips']} flips / {r['total']} codes{mm}{extra}")
    tr.write_line(f"TOTAL: {flips} flips / {total} codes compared")
