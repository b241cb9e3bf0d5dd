"""TEST INFRASTRUCTURE ONLY — CPU restatement of the ViT-2SPN dual-stream SSP step.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker (or the timed CPU baseline), never as the thing shipped.
"""
