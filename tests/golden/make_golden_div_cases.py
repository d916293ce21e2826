"""Cases and the seeded loss shared by tests/golden/make_golden_div.py (build container) and tests/test_gpu_divergence.py (GPU box);
imports nothing from the upstream tree."""
import scenes

EXPECTED_GRAD_CASES = [("toy_world", 2), ("toy_world", 0), ("tennis_dense", 1)]


def expected_loss(name, k, exp, opacity):
    c1 = scenes.cotangent(f"expected/{name}/{k}", exp.shape).to(exp.device)
    c2 = scenes.cotangent(f"expected_opacity/{name}/{k}", opacity.shape).to(exp.device)
    return (c1 * exp).sum() + (c2 * opacity).sum()
