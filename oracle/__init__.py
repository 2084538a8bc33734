"""CPU oracle for AffectGPT's multimodal encode-and-fuse path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``affectgpt_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs do, and there only as the checker / the timed CPU arm.

Parity status: **unpinned at the HuggingFace boundary** -- the reference ships
no numerical test, golden vector or known-answer fixture for this path
(finding F10); the arithmetic lives in the un-vendored
``transformers`` dependency (pinned 4.49.0 by the reference, 5.5.0 in this
image).  The oracle therefore is: the HuggingFace ``CLIPVisionModelWithProjection``
/ ``HubertModel`` modules themselves (random-init, fp32, CPU) wrapped by a
restatement of the reference's thin wrappers and fusion heads, each function
citing the reference file:line it follows.  The integer contracts (frame
indices, audio segment starts, token splice positions) ARE pinned: against the
known answers derived in SURVEY.md section 8a (a1, a4) and the reference's demo
fixture sizes.
"""
