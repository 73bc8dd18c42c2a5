"""Batched decode (emx_decode_batch_step, BASELINE.json configs[4]: 8 sequences per GPU, mixed max_new_tokens) against
(1) the one-sequence-per-launch kernel run on each sequence alone and (2) the torch-eager oracle on the same GPU.

The reference cannot batch cached generation (bs == 1 asserts, modeling_prismatic.py:326, :460-463), so its result for a batch IS
its bs=1 result per sequence; that is the bar: greedy ids bit-equal to each sequence's own bs=1 run and to the oracle, per-step
logits (teacher-forced) within 1e-2 of max|logit| of the oracle (2e-2 on un-scripted O(1) logits at full size, see
test_gpu_e2e.py::test_full_size_unscripted_head_teacher_forced)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _rel_err(got, want):
    got, want = got.float(), want.float()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()


def _oracle_attn():
    try:
        import flash_attn  # noqa: F401

        return "flash_attention_2"
    except Exception:
        return "sdpa"


def _prompts(cfg, lens, seed):
    rng = np.random.default_rng(seed)
    V = cfg.text_config.vocab_size
    return [torch.tensor([[1] + rng.integers(3, V - 64, n - 1).tolist()], dtype=torch.long, device="cuda") for n in lens]


def _pixels(B, seed):
    return torch.randn((B, 6, 224, 224), generator=torch.Generator(device="cuda").manual_seed(seed), device="cuda").to(BF)


def _oracle_runs(cfg, sd, ids, pv, n_new, attn):
    """Each sequence through the oracle alone (what the reference would do): greedy ids and per-step logits."""
    from oracle.model import OracleVLA

    oracle = OracleVLA.from_state_dict(cfg, sd, device="cuda", dtype=BF, attn_implementation=attn)
    out = []
    for b, x in enumerate(ids):
        ids_o, logits_o = oracle.generate(x, pv[b : b + 1], n_new[b], eos_token_id=None, return_logits=True)
        out.append((ids_o[0, x.shape[1] :].tolist(), logits_o))
    del oracle
    torch.cuda.empty_cache()
    return out


def _check_batch_vs_single_and_oracle(cfg, sd, lens, n_new, seed, attn, tol, max_context=1024):
    from emmax_b200 import OpenVLAForActionPrediction

    B = len(lens)
    ids, pv = _prompts(cfg, lens, seed), _pixels(B, seed)
    want = _oracle_runs(cfg, sd, ids, pv, n_new, attn)
    forced = [w[0] for w in want]
    model = OpenVLAForActionPrediction(cfg, dict(sd), max_batch=8, max_context=max_context).to("cuda")
    eng = model.engine
    same = len(set(lens)) == 1
    batch_in = torch.cat(ids, dim=0) if same else ids
    new_b, logits_b = eng.generate_batch(batch_in, pv, n_new, eos_token_id=None, return_logits=True, forced=forced)
    torch.cuda.synchronize()
    logits_b = logits_b.cpu()
    for b in range(B):
        assert new_b[b].numel() == n_new[b], f"sequence {b}: {new_b[b].numel()} tokens, limit {n_new[b]}"
        # (1) the bs=1 kernel on this sequence alone, same teacher forcing
        _, logits_1 = eng.generate(ids[b], pv[b : b + 1], n_new[b], eos_token_id=None, return_logits=True, forced=forced[b])
        lb, l1, lo = logits_b[: n_new[b], b], logits_1.cpu(), want[b][1]
        e1 = _rel_err(lb, l1)
        eo = _rel_err(lb, lo)
        assert e1 < tol, f"sequence {b} (prompt {lens[b]}, {n_new[b]} new): batched vs bs=1 kernel logits rel err {e1:.4g} (tolerance {tol})"
        assert eo < tol, f"sequence {b}: batched kernel vs oracle logits rel err {eo:.4g} (tolerance {tol})"
        # a sequence must not see its neighbours: the argmax agrees with the oracle wherever the oracle's margin is clear
        top2 = lo.float().topk(2, dim=-1).values
        clear = (top2[:, 0] - top2[:, 1]) > 4 * tol * float(lo.abs().max())
        assert bool((lb.argmax(-1)[clear] == lo.argmax(-1)[clear]).all()), f"sequence {b}: argmax differs at a clear oracle margin"
    return model


def test_tiny_batch_mixed_lengths_and_limits_vs_single_and_oracle():
    """Toy widths, un-scripted head, B = 5 with different prompt lengths (prefilled one by one) and different token limits: contexts
    279..~330 = 5-6 pages = 3 page-pair items per (sequence, head) row spread over different CTAs (multi-segment combine)."""
    from emmax_b200 import tiny_config
    from emmax_b200.synthetic import make_state_dict

    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=3, device="cpu")
    _check_batch_vs_single_and_oracle(cfg, sd, lens=[23, 31, 17, 40, 23], n_new=[12, 30, 21, 30, 5], seed=5, attn="sdpa", tol=1e-2)


def test_tiny_batch_of_eight_same_length_long_context():
    """B = 8 through ONE batched prefill; 70 new tokens cross a page boundary (context 290 -> 360: 5 -> 6 pages)."""
    from emmax_b200 import tiny_config
    from emmax_b200.synthetic import make_state_dict

    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=4, device="cpu")
    _check_batch_vs_single_and_oracle(cfg, sd, lens=[34] * 8, n_new=[70, 70, 33, 70, 70, 9, 70, 70], seed=6, attn="sdpa", tol=1e-2)


def test_tiny_batch_scripted_ids_eos_and_limits():
    """Scripted head (planted greedy continuation, margins >> bf16 noise), free-running (no teacher forcing): every sequence's ids are
    bit-equal to the script; the script ends in EOS, which stops each sequence on the device; limits below the script length cut it."""
    from emmax_b200 import OpenVLAForActionPrediction, SyntheticLlamaTokenizer, tiny_config
    from emmax_b200.synthetic import default_script, make_state_dict

    cfg = tiny_config()
    tok = SyntheticLlamaTokenizer()
    script = default_script(tok, 48, seed=2)
    prev = 77
    sd = make_state_dict(cfg, seed=9, device="cpu", script=script, script_prev=prev)
    model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
    lens = [12, 20, 12, 33, 12, 12]
    ids = _prompts(cfg, lens, 1)
    for x in ids:
        x[0, -1] = prev  # every prompt ends in the script's predecessor
    pv = _pixels(len(lens), 2)
    limits = [64, 64, 10, 64, 48, 47]
    new, _ = model.engine.generate_batch(ids, pv, limits, eos_token_id=tok.eos_token_id)
    for b, lim in enumerate(limits):
        want = script[: min(lim, len(script))]
        assert new[b].cpu().tolist() == want, f"sequence {b} (limit {lim}): ids differ from the script"
    # and each one equals its own bs=1 run
    for b in (1, 3):
        one, _ = model.engine.generate(ids[b], pv[b : b + 1], limits[b], eos_token_id=tok.eos_token_id)
        assert torch.equal(one, new[b])


def test_full_size_batch_of_eight_vs_single_and_oracle():
    """Full Emma-X shapes (Llama-2-7B widths), un-scripted head, B = 8, mixed limits: every sequence against its own bs=1 run and the
    flash-attn oracle."""
    from emmax_b200 import emma_x_config
    from emmax_b200.synthetic import make_state_dict

    cfg = emma_x_config()
    sd = make_state_dict(cfg, seed=7, device="cuda")
    _check_batch_vs_single_and_oracle(cfg, sd, lens=[40] * 8, n_new=[24, 12, 24, 24, 6, 24, 18, 24], seed=8, attn=_oracle_attn(), tol=2e-2)


def test_full_size_batch_scripted_512_tokens_mixed_limits():
    """BASELINE.json configs[4] at one GPU's share: 8 sequences, half with 128 and half with 512 new tokens, full size, scripted head,
    free-running: ids bit-equal to the script for every sequence (contexts 296 -> 808)."""
    from emmax_b200 import OpenVLAForActionPrediction, SyntheticLlamaTokenizer, emma_x_config
    from emmax_b200.synthetic import default_script, make_state_dict

    cfg = emma_x_config()
    tok = SyntheticLlamaTokenizer()
    script = default_script(tok, 512, seed=0)
    ids = _prompts(cfg, [40] * 8, 1234)
    prev = int(ids[0][0, -1])
    for x in ids:
        x[0, -1] = prev
    sd = make_state_dict(cfg, seed=0, device="cuda", script=script, script_prev=prev)
    model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
    pv = _pixels(8, 3)
    limits = [128, 512, 128, 512, 128, 512, 128, 512]
    new, _ = model.engine.generate_batch(torch.cat(ids, 0), pv, limits, eos_token_id=None)
    for b, lim in enumerate(limits):
        assert new[b].cpu().tolist() == script[:lim], f"sequence {b} (limit {lim}): ids differ from the script"


def test_simpler_policy_on_the_engine_single_and_batched():
    """the SimplerEnv policy wrapper driving the REAL engine. `OpenVLAInference.step` (openvla_model.py:72-145) on one
    environment, and `BatchedOpenVLAInference.step` over 5 environments (different frames and tasks) in one batched decode: every
    environment's raw action and post-processed action equal its own single-environment step, bit for bit."""
    from emmax_b200 import AutoProcessor, BatchedOpenVLAInference, OpenVLAForActionPrediction, OpenVLAInference, tiny_config
    from emmax_b200.synthetic import make_state_dict, predict_action_chain

    cfg = tiny_config()
    stats = cfg.norm_stats["synthetic"]
    cfg.norm_stats = {"bridge_orig": stats}
    # plant the predict_action chain (29871 -> 7 action ids) so the greedy action tokens are well separated from bf16 noise
    prev0, chain = predict_action_chain([])
    sd = make_state_dict(cfg, seed=21, device="cpu", extra_chains=[(prev0, chain)])
    model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
    proc = AutoProcessor.from_pretrained(None)
    rng = np.random.default_rng(4)
    n_env = 5
    frames = [rng.integers(0, 256, (256, 320, 3), dtype=np.uint8) for _ in range(n_env)]
    tasks = ["put carrot in pot", "open the drawer", "put carrot in pot", "stack the green block on the yellow block", "close drawer"]
    kw = dict(policy_setup="widowx_bridge", vla=model, processor=proc, device="cuda")
    batched = BatchedOpenVLAInference(n_env, **kw)
    singles = [OpenVLAInference(**kw) for _ in range(n_env)]
    for step in range(2):
        got = batched.step(frames, tasks)
        for e in range(n_env):
            raw, act = singles[e].step(frames[e], tasks[e])
            for k in raw:
                assert np.array_equal(got[e][0][k], raw[k]), (step, e, k)
            for k in act:
                assert np.array_equal(np.asarray(got[e][1][k]), np.asarray(act[k])), (step, e, k)
        frames = [np.roll(f, 7, axis=1) for f in frames]
    # the planted chain really is what came out: the un-normalised action of the chain's tokens
    want = model.detokenize_on_device(torch.tensor(chain, dtype=torch.int32, device="cuda"))[1].cpu().numpy()
    raw, _ = singles[0].step(frames[0], tasks[0])
    assert np.array_equal(np.concatenate([raw["world_vector"], raw["rotation_delta"], raw["open_gripper"]]), want)


def test_tiny_continuous_batching_stream_equals_single_runs():
    """Engine.serve: 19 requests (different prompt lengths, different token limits) streamed through the 8 sequence slots — a slot is
    refilled the moment its sequence reaches its limit, same-length neighbours share one batched prefill into whatever slots are free.
    Scripted head (argmax margins far above bf16 noise, so greedy ids are a bit-exact check: the flash-decoding split of a row depends on
    its neighbours' contexts and moves low-order bits): every request's ids equal the script AND its own bs=1 `generate`; neighbours,
    slot placement and refills must not matter."""
    from emmax_b200 import OpenVLAForActionPrediction, SyntheticLlamaTokenizer, tiny_config
    from emmax_b200.synthetic import default_script, make_state_dict

    cfg = tiny_config()
    script = default_script(SyntheticLlamaTokenizer(), 30, seed=6)[:-1]  # (without the closing EOS: limits alone end the sequences)
    prev = 83
    sd = make_state_dict(cfg, seed=11, device="cpu", script=script, script_prev=prev)
    model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
    eng = model.engine
    lens = [23, 23, 23, 31, 17, 17, 40, 23, 23, 12, 12, 12, 12, 29, 23, 23, 18, 18, 33]
    limits = [9, 29, 5, 17, 29, 1, 12, 25, 7, 29, 3, 14, 22, 6, 29, 11, 2, 19, 8]
    ids, pv = _prompts(cfg, lens, 13), _pixels(len(lens), 14)
    for x in ids:
        x[0, -1] = prev
    got = eng.serve([(ids[i], pv[i : i + 1], limits[i]) for i in range(len(lens))], eos_token_id=None)
    torch.cuda.synchronize()
    assert eng.last_serve["launches"] < sum(limits) // 4, "slots were not shared"
    for i in range(len(lens)):
        assert got[i].cpu().tolist() == script[: limits[i]], f"request {i} (prompt {lens[i]}, limit {limits[i]}): ids differ from the script"
    for i in (1, 4, 6, 13):
        one, _ = eng.generate(ids[i], pv[i : i + 1], limits[i], eos_token_id=None)
        assert torch.equal(got[i], one), f"request {i}: stream result differs from its bs=1 run"
    # longest-first admission (an offline batch with known limits): same ids per request, in request order, in no more launches
    fifo_launches = eng.last_serve["launches"]
    lpt = eng.serve([(ids[i], pv[i : i + 1], limits[i]) for i in range(len(lens))], eos_token_id=None, order="longest_first")
    torch.cuda.synchronize()
    for i in range(len(lens)):
        assert torch.equal(lpt[i], got[i]), f"request {i}: longest-first result differs from the FIFO stream's"
    assert eng.last_serve["launches"] <= fifo_launches, (eng.last_serve["launches"], fifo_launches)


def test_tiny_continuous_batching_scripted_eos():
    """Scripted head ending in EOS: with EOS active the stream retires sequences at the lagged poll; ids bit-equal to the script (cut at
    each limit), through the model-level `generate_batch(..., continuous=True)`."""
    from emmax_b200 import OpenVLAForActionPrediction, SyntheticLlamaTokenizer, tiny_config
    from emmax_b200.synthetic import default_script, make_state_dict

    cfg = tiny_config()
    tok = SyntheticLlamaTokenizer()
    script = default_script(tok, 40, seed=5)
    prev = 91
    sd = make_state_dict(cfg, seed=12, device="cpu", script=script, script_prev=prev)
    model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
    lens = [12, 20, 12, 33, 12, 12, 20, 20, 15, 15, 15, 12]
    limits = [64, 64, 10, 64, 40, 39, 64, 7, 64, 64, 21, 64]
    ids = _prompts(cfg, lens, 3)
    for x in ids:
        x[0, -1] = prev
    pv = _pixels(len(lens), 4)
    out = model.generate_batch(ids, [pv[i : i + 1] for i in range(len(lens))], limits, eos_token_id=tok.eos_token_id, continuous=True)
    for i, lim in enumerate(limits):
        want = script[: min(lim, len(script))]
        assert out[i][0, lens[i] :].cpu().tolist() == want, f"request {i} (limit {lim}): ids differ from the script"
