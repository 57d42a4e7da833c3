"""CPU oracle for the Genima per-step inference hot path.

TEST INFRASTRUCTURE ONLY.  Everything under oracle/ is a plain fp32 PyTorch-on-CPU restatement of the arithmetic the
reference reaches through diffusers 0.29.0 / RoboBase / OpenAI-CLIP / torchvision (none of which is vendored in the
reference repository or installable offline — SURVEY.md §8c).  It may be imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and only as the checker or the timed
CPU baseline; the product path (genima_b200/) never imports it and fails loudly without its CUDA library.

PARITY UNPINNED for the diffusers / RoboBase graphs: the reference ships no tests, golden vectors or weights for this
path, and the upstream packages are absent, so those graphs are restated from their published architecture
(SURVEY.md Appendices A-F) and cross-checked structurally (exact parameter counts, state-dict key names).  Pinned
pieces: the CLIP text towers against transformers' CLIPTextModel, the ResNet-18 trunk against torchvision, multi-head
attention against torch.nn.MultiheadAttention, tile/untile against the reference's own controller/utils/misc.py
(imported in this container to generate tests/golden/), and the Euler scheduler tables against closed-form values.
"""
