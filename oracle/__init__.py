"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms (``liboracle.so``) and ctypes doors onto the
reference's own code compiled unmodified (``_ref/libref_cpu.so``, ``_ref/libref_cuda.so``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this package.  Nothing under ``geometrics_b200/`` does: the
product path has no CPU fallback and fails loudly when its CUDA library is missing.
"""
from .refload import (  # noqa: F401
    Oracle,
    RefCpu,
    RefCuda,
    build_oracle,
    have_ref_cpu,
    have_ref_cuda,
)
