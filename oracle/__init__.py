"""CPU oracle for the paged-KV decode/attention path (TEST INFRASTRUCTURE, not product).

A restatement, in plain PyTorch-on-CPU, of the reference's own CPU/eager implementation of
the path (``nanovllm/layers/attention_torch_native.py`` + the eager layers it sits between).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package; nothing under ``nano-vllm-ascend_b200/`` does.

Parity status: PINNED. The reference publishes no golden vectors for this path,
so the pin is made with outputs of the reference itself: ``oracle/gen_golden.py`` imports the
unmodified reference modules from /root/reference (namespace stub), runs them on seeded inputs
and stores the results in ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` asserts that
every function here reproduces those outputs bit for bit on CPU.
"""
