"""CPU oracle for AffectGPT's multimodal encode-and-fuse path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``affectgpt_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
legs (``cpu_baseline`` and ``--impl reference``) do, and there only as the
checker / the timed CPU arm.

What it is: the reference's own thin code (encoder wrappers, fusion heads,
Q-Former fusion variants, samplers, audio segmentation, token splice) restated
function by function, each citing the reference file:line it follows, on top
of the HuggingFace ``CLIPVisionModelWithProjection`` / ``HubertModel`` modules
(random-init, fp32, CPU) - the un-vendored ``transformers`` dependency where
the reference's heavy arithmetic lives (pinned 4.49.0 by the reference, 5.5.0
in this image).

Parity status: **pinned against the reference itself, run in the build
container.**  The reference ships no golden vectors and cannot be
imported as a package here (omegaconf, decord, pytorchvideo, peft are not
installed), so ``tests/golden/ref_loader.py`` executes the
reference's UNMODIFIED source files (``encoder.py``, ``affectgpt.py``,
``Qformer.py``, ``video_processor.py``, ``ImageBind/data.py``) with stub modules
for the missing third-party imports, and ``tests/golden/make_reference_golden.py``
records their outputs as fixtures (``tests/golden/reference_*.{json,npz}``):

* frame / face index rules, the emotion-peak sampler incl. its JSON loader, the
  head/tail sampler under a seeded ``random``, audio slice bounds - bit-exact;
* ``cv2.resize`` of face crops through the reference's ``load_face`` - bit-exact;
* fusion heads (attention / mean / pre-extracted twins) and all three
  Q-Former variants from the reference's ``AffectGPT`` methods over its own
  ``Qformer.py`` - the oracle matches to 0 / 1e-6;
* the ``CLIP_VIT_LARGE`` / ``HUBERT_LARGE`` wrapper classes (constructor from a
  checkpoint directory + forward) - the oracle matches to 1e-6 / 0.

Two things remain outside that pin and are stated wherever they matter:
the pytorchvideo clip sampler (package absent; restated from its published
0.1.5 source in ``oracle/sampling.py``) and two API-drift shims between
transformers 4.49 and 5.5 (``AutoFeatureExtractor`` for CLIP, the return type of
``get_image_features``), neither of which touches the arithmetic.
"""
