"""Sibling Video-ViT architectures the reference fine-tunes next to its own model (other_models/ in the reference):
MVD (3-D sin-cos position table, optional class token) and UMT (interpolated sinusoid table, tubelet 1).  Both run the
same sm_100a kernels through libstad.so; only the host-side model description differs."""
