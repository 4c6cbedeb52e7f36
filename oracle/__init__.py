"""CPU oracle for the WalkGPT pixel-grounding forward path (Path A).

TEST INFRASTRUCTURE ONLY.  Nothing under ``walkgpt_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and only as the checker / reported CPU baseline.

``oracle.path_a`` is a plain fp32 PyTorch restatement of the reference's module arithmetic,
written functionally over ``state_dict`` tensors that use the reference's parameter names.
Parity status: PINNED by execution of the reference itself -- ``oracle/make_golden.py``
imports the reference modules from ``/root/reference`` (and HF ``transformers`` CLIP, which is
where the reference's CLIP arithmetic lives), loads the same seeded weights with
``load_state_dict(strict=True)``, and writes the outputs to ``tests/golden/``; the CPU test-suite
checks this restatement against those fixtures.  The reference ships no tests or golden
vectors of its own.  The relative-depth head has NO reference implementation
(SURVEY.md §0): ``oracle.path_a.depth_head`` is this repo's own definition -> "parity unpinned".
"""
