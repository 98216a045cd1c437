"""Host half of the temperature ladder (upstream transcribe.py::decode_with_fallback, batched in
``WhisperB200._decode_with_fallback``) with the device decode replaced by scripted results.  CPU only."""
import pytest
import torch

from whisperjav_b200 import model as M


def _res(tokens, avg, cr, ns, t):
    return M.DecodingResult(tokens=list(tokens), text="x" * len(tokens), avg_logprob=avg, no_speech_prob=ns, temperature=t,
                            compression_ratio=cr, language="ja", sum_logprob=avg * (len(tokens) + 1))


class _Scripted(M.WhisperB200):
    """No device: ``_decode_with_prompts`` pops scripted per-window results and records what it was asked."""

    def __init__(self, script):
        self.script = script      # {(window id, temperature, draw): DecodingResult}
        self.calls = []
        self._sample_calls = 0
        self.device = "cpu"
        self.draw = {}

    def _decode_with_prompts(self, xa, prompts, temperature, language, task, decode_options, seed=0):
        ids = [int(v) for v in xa[:, 0].tolist()]
        self.calls.append((temperature, ids, seed))
        self.sample_lens = getattr(self, "sample_lens", []) + [decode_options.get("sample_len")]
        out = []
        for i in ids:
            k = self.draw.get((i, temperature), 0)
            self.draw[(i, temperature)] = k + 1
            r = self.script[(i, temperature, k)]
            cap = decode_options.get("sample_len")
            if cap is not None and len(r.tokens) >= cap:   # the device stops at the cap: no EOT yet
                r = M.DecodingResult(tokens=r.tokens[:cap], text=r.text[:cap], avg_logprob=r.avg_logprob, no_speech_prob=r.no_speech_prob,
                                     temperature=r.temperature, compression_ratio=r.compression_ratio, sum_logprob=r.sum_logprob,
                                     complete=False, steps_needed=cap)
            out.append(r)
        return out


def _oracle_ladder(script, i, temps, best_of, cr_thr, lp_thr, ns_thr):
    """decode_with_fallback for one window, best_of candidates ranked by sum_logprob / length at t > 0."""
    res = None
    for t in temps:
        n = best_of if (t > 0 and best_of > 1) else 1
        cands = [script[(i, t, k)] for k in range(n)]
        res = max(cands, key=lambda r: r.sum_logprob / max(len(r.tokens), 1)) if n > 1 else cands[0]
        needs = False
        if cr_thr is not None and res.compression_ratio > cr_thr:
            needs = True
        if lp_thr is not None and res.avg_logprob < lp_thr:
            needs = True
        if ns_thr is not None and res.no_speech_prob > ns_thr and lp_thr is not None and res.avg_logprob < lp_thr:
            needs = False
        if not needs:
            break
    return res


@pytest.mark.parametrize("best_of", [1, 3])
def test_ladder_redecodes_only_the_failing_windows(best_of):
    temps = [0.0, 0.4, 0.8]
    good = lambda t: _res([5, 6, 7], -0.3, 1.2, 0.1, t)          # noqa: E731
    repetitive = lambda t: _res([5] * 9, -0.2, 3.1, 0.1, t)      # noqa: E731  compression ratio too high
    unsure = lambda t: _res([8, 9], -1.7, 1.1, 0.2, t)           # noqa: E731  avg_logprob too low
    silent = lambda t: _res([4], -1.9, 1.0, 0.95, t)             # noqa: E731  low logprob but judged silent: accepted
    script = {}
    for t in temps:
        for k in range(3):
            script[(0, t, k)] = good(t)
            script[(1, t, k)] = repetitive(t) if t < 0.8 else _res([5, 6, 8, 9][: 2 + k], -0.5 + 0.1 * k, 1.4, 0.1, t)
            script[(2, t, k)] = unsure(t) if t == 0.0 else _res([8, 9, 10 + k], -0.9 + 0.2 * k, 1.1, 0.2, t)
            script[(3, t, k)] = silent(t)
            script[(4, t, k)] = unsure(t)                          # never recovers: the last temperature's result stands
    m = _Scripted(script)
    xa = torch.arange(5, dtype=torch.float32).view(5, 1)
    out = m._decode_with_fallback(xa, [[]] * 5, temps, best_of, "ja", "transcribe", {}, 2.4, -1.0, 0.6)
    for i in range(5):
        want = _oracle_ladder(script, i, temps, best_of, 2.4, -1.0, 0.6)
        assert out[i].tokens == want.tokens and out[i].temperature == want.temperature and out[i].avg_logprob == want.avg_logprob, i
    # batches shrink: t = 0 sees all windows once, later temperatures only the failing ones, best_of draws each
    by_t = {}
    for t, ids, seed in m.calls:
        by_t.setdefault(t, []).append(ids)
    assert by_t[0.0] == [[0, 1, 2, 3, 4]]
    assert by_t[0.4] == [[1, 2, 4]] * best_of
    assert by_t[0.8] == [[1, 4]] * best_of
    assert len({seed for _, _, seed in m.calls}) == len(m.calls)   # every sampled pass gets its own seed


def test_thresholds_can_be_disabled():
    script = {(0, 0.0, 0): _res([5] * 9, -3.0, 9.9, 0.0, 0.0), (0, 0.5, 0): _res([1], -0.1, 1.0, 0.0, 0.5)}
    m = _Scripted(script)
    out = m._decode_with_fallback(torch.zeros(1, 1), [[]], [0.0, 0.5], 1, "ja", "transcribe", {}, None, None, None)
    assert out[0].temperature == 0.0 and [c[0] for c in m.calls] == [0.0]


def test_step_cap_keeps_cut_off_windows_out_of_the_ladder():
    """A first-tier pass with a step cap (StepCapPlanner): windows that had not ended come back incomplete and are neither accepted
    nor sent up the temperature ladder; the others behave exactly as without the cap; later temperatures run uncapped."""
    temps = [0.0, 0.4]
    script = {}
    for t in temps:
        script[(0, t, 0)] = _res([5, 6, 7], -0.3, 1.2, 0.1, t)                        # short and fine
        script[(1, t, 0)] = _res(list(range(100, 140)), -0.3, 1.2, 0.1, t)            # 40 tokens: cut off by a cap of 16
        script[(2, t, 0)] = _res([8, 9], -1.7, 1.1, 0.2, t) if t == 0.0 else _res(list(range(30)), -0.5, 1.1, 0.2, t)   # short, falls back
    m = _Scripted(script)
    xa = torch.arange(3, dtype=torch.float32).view(3, 1)
    out = m._decode_with_fallback(xa, [[]] * 3, temps, 1, "ja", "transcribe", {}, 2.4, -1.0, 0.6, step_cap=16)
    assert out[0].complete and out[0].tokens == [5, 6, 7]
    assert not out[1].complete and len(out[1].tokens) == 16
    assert out[2].complete and out[2].temperature == 0.4 and len(out[2].tokens) == 30   # the T = 0.4 pass was not capped
    assert [c[:2] for c in m.calls] == [(0.0, [0, 1, 2]), (0.4, [2])] and m.sample_lens == [16, None]
    # uncapped: the same windows, window 1 whole
    m2 = _Scripted(script)
    out2 = m2._decode_with_fallback(xa, [[]] * 3, temps, 1, "ja", "transcribe", {}, 2.4, -1.0, 0.6)
    assert [o.tokens for o in out2] == [[5, 6, 7], list(range(100, 140)), list(range(30))] and out2[0].tokens == out[0].tokens


def test_step_cap_planner_picks_the_cheapest_ladder_and_stays_off_when_it_does_not_pay():
    mk = lambda n, done=True: M.DecodingResult(tokens=[1] * n, complete=done, steps_needed=n)   # noqa: E731
    p = M.StepCapPlanner(full=224, min_obs=32)
    assert p.ladder() == [] and p.cap() is None             # nothing seen yet: first pass runs uncapped
    p.observe([mk(20 + (k % 10)) for k in range(60)] + [mk(224) for _ in range(4)])   # 6 % stragglers at the limit
    assert p.ladder() == [32] and p.cap() == 32 and p.cap(1) is None   # 32 + 232 * 4/64 = 46.5 steps per pass instead of 224; no middle tier pays
    q = M.StepCapPlanner(full=224, min_obs=32)
    q.observe([mk(200 + (k % 20)) for k in range(64)])      # everything long: capping buys nothing
    assert q.ladder() == []
    u = M.StepCapPlanner(full=224, min_obs=32)
    u.observe([mk(30) for _ in range(64)])                  # uniform lengths: the pass already ends at 30
    assert u.ladder() == []
    # a long middle: 70 % short, 25 % around 80-90, 5 % never end -> two caps
    t = M.StepCapPlanner(full=224, min_obs=32)
    t.observe([mk(10 + (k % 6)) for k in range(90)] + [mk(80 + (k % 10)) for k in range(32)] + [mk(224) for _ in range(6)])
    lad = t.ladder()
    assert len(lad) == 2 and lad[0] == 16 and lad[1] == 96, lad
    o = t.PASS_OVERHEAD
    one = min(c + (224 + o) * t._share_longer(c) for c in t.CANDIDATES)
    two = lad[0] + t._share_longer(lad[0]) * (lad[1] + o) + t._share_longer(lad[1]) * (224 + o)
    assert two < 0.97 * one
    # windows cut off count as full until a later tier tells their length
    r = M.StepCapPlanner(full=224, min_obs=32)
    r.observe([mk(10) for _ in range(48)] + [mk(16, done=False) for _ in range(16)])
    assert r.cut == 16 and r.obs.count(224) == 16
    r.resolve([mk(60) for _ in range(10)] + [mk(96, done=False) for _ in range(2)])
    assert r.cut == 6 and r.obs.count(224) == 6 and r.obs.count(60) == 10


def test_tier_scheduler_finishes_every_window_once_with_its_full_result():
    """Scripted device: window i needs L[i] steps; a pass costs max(min(L, cap)) steps.  Every window must be finished exactly once with
    the tokens an uncapped decode gives, the pools must drain, and the run must cost fewer steps than one-tier decoding."""
    import random
    rnd = random.Random(3)
    n, B, full = 640, 64, 224
    L = [rnd.choice([8, 9, 11, 14, 20, 26]) if rnd.random() < 0.72 else (rnd.randint(60, 100) if rnd.random() < 0.8 else full) for _ in range(n)]
    cost = {"steps": 0, "passes": 0}
    done = {}

    def decode(rows, idx, cap):
        assert rows.shape[0] == len(idx) and [int(v) for v in rows[:, 0].tolist()] == idx   # the encoder rows travel with their windows
        lim = full if cap is None else cap
        cost["steps"] += max(min(L[i], lim) for i in idx)
        cost["passes"] += 1
        return [M.DecodingResult(tokens=[i] * min(L[i] - 1, lim), complete=(L[i] <= lim) if cap is not None else True, steps_needed=min(L[i], lim))
                for i in idx]

    def finish(idx, rows, sizes, results):
        for i, z, r in zip(idx, sizes, results):
            assert i not in done and z == 100 + i and r.complete
            done[i] = r.tokens

    stats = {}
    sched = M.TierScheduler(M.StepCapPlanner(full, min_obs=32), B, decode, finish, stats)
    for c0 in range(0, n, B):
        idx = list(range(c0, c0 + B))
        sched.submit(idx, torch.tensor(idx, dtype=torch.float32).view(-1, 1).repeat(1, 3), [100 + i for i in idx])
    sched.drain()
    assert sorted(done) == list(range(n)) and all(done[i] == [i] * (L[i] - 1) for i in range(n))
    assert all(not p for p in sched.pools.values())
    one_tier = sum(max(L[c0: c0 + B]) for c0 in range(0, n, B))
    assert stats["windows_redecoded"] > 0 and cost["steps"] < 0.7 * one_tier, (cost, one_tier, stats)
    # no planner: plain passes
    done.clear()
    cost.update(steps=0, passes=0)
    plain = M.TierScheduler(None, B, decode, finish, {})
    for c0 in range(0, n, B):
        idx = list(range(c0, c0 + B))
        plain.submit(idx, torch.tensor(idx, dtype=torch.float32).view(-1, 1).repeat(1, 3), [100 + i for i in idx])
    plain.drain()
    assert cost["steps"] == one_tier and cost["passes"] == n // B and sorted(done) == list(range(n))
