"""CPU oracle for the IoU-aware RetinaNet inference hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package -- always as the checker or the timed CPU baseline, never as part of the
product path (the product path raises if the CUDA library is missing).
"""
