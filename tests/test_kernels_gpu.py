"""GPU parity of every kernel behind the C ABI against PyTorch fp32 (see tests/kernel_checks.py)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [
    "cast", "row_stats", "layernorm", "pool_head", "gemm_plain_small", "gemm_plain_k", "gemm_plain", "gemm_bn",
    "gemm_resid", "gemm_ln", "gemm_ln_gelu", "gemm_stats", "attention_small", "attention_tail", "attention", "attention_ragged", "attention_persistent", "patch_embed",
    "patch_embed_frames", "patch_embed_masked", "decoder_assemble", "tail_rows", "normalize_u8", "gemm_e2d_shapes", "gemm_pair", "resize_cubic",
    "rows_norm_head", "prepend_cls", "attention_siblings", "patch_embed_siblings", "attention_outliers",
    "gemm_integer_exact", "attention_full_size", "attention_edges", "gemm_edges",
])
def test_kernel(name):
    from tests.kernel_checks import CHECKS
    CHECKS[name]()
