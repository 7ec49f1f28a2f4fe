"""Generate the golden fixtures in this directory by EXECUTING the reference (raphaelsty/mkb).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports ``river.stats`` for two trivial accumulators; ``river`` is not
installed, so a throw-away shim (Mean / RollingMean) is written to a temp dir and put on
``sys.path``.  Nothing from the reference is copied into the repo: only the *outputs* of
running it (scores, losses, autograd gradients, sampler draws, filter lists, ranks) are
stored as small ``.npz`` files next to this script.
"""
import collections
import os
import sys
import tempfile
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _install_shim():
    d = tempfile.mkdtemp(prefix="river_shim_")
    os.makedirs(os.path.join(d, "river"))
    with open(os.path.join(d, "river", "__init__.py"), "w") as f:
        f.write("from . import stats\n")
    with open(os.path.join(d, "river", "stats.py"), "w") as f:
        f.write(
            "import collections\n"
            "class Mean:\n"
            "    def __init__(self): self.n = 0; self.m = 0.0\n"
            "    def update(self, x): self.n += 1; self.m += (x - self.m) / self.n; return self\n"
            "    def get(self): return self.m\n"
            "class RollingMean:\n"
            "    def __init__(self, window_size): self.d = collections.deque(maxlen=window_size)\n"
            "    def update(self, x): self.d.append(x); return self\n"
            "    def get(self): return sum(self.d) / len(self.d) if self.d else 0.0\n"
        )
    sys.path.insert(0, d)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True


_install_shim()
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from mkb import datasets, evaluation, losses, models, sampling  # noqa: E402
from mkb.datasets import base as ref_base  # noqa: E402

MODEL_GAMMA = {"TransE": 6.0, "DistMult": 9.0, "ComplEx": 9.0, "RotatE": 9.0}


def _np(t):
    return t.detach().cpu().numpy()


def gen_step_cases():
    """scores / loss / autograd grads for 4 models x 2 modes x {aligned, unaligned} dims."""
    out = {}
    rng = np.random.RandomState(7)
    N, R, B, K = 37, 5, 6, 7
    ent_map = {f"e{i}": i for i in range(N)}
    rel_map = {f"r{i}": i for i in range(R)}
    for name, gamma in MODEL_GAMMA.items():
        for D in (8, 5):
            torch.manual_seed(1000 + D)
            model = getattr(models, name)(hidden_dim=D, entities=ent_map, relations=rel_map, gamma=gamma)
            # widen the init a little so scores are not all ~gamma and softmax weights differ
            with torch.no_grad():
                model.entity_embedding.mul_(3.0)
                model.relation_embedding.mul_(3.0)
            for mode in ("tail-batch", "head-batch"):
                sample = np.stack(
                    [rng.randint(N, size=B), rng.randint(R, size=B), rng.randint(N, size=B)], axis=1
                ).astype(np.int64)
                neg = rng.randint(N, size=(B, K)).astype(np.int64)
                neg[0, 1] = neg[0, 0]  # duplicate negative inside a row
                neg[1, 0] = sample[1, 0]  # negative equal to the positive's head
                neg[2, 0] = sample[2, 2]  # negative equal to the positive's tail
                sample[3] = sample[2]  # duplicated positive
                weight = rng.uniform(0.1, 0.5, size=B).astype(np.float32)
                key = f"{name}_D{D}_{mode}"
                for tag, mdl in (("f32", model), ("f64", None)):
                    if mdl is None:
                        import copy

                        mdl = copy.deepcopy(model).double()
                    mdl.zero_grad()
                    s_t = torch.from_numpy(sample)
                    n_t = torch.from_numpy(neg)
                    w_t = torch.from_numpy(weight)
                    if tag == "f64":
                        w_t = w_t.double()
                    pos = mdl(s_t)
                    ngs = mdl(s_t, n_t, mode)
                    loss = losses.Adversarial(alpha=0.5)(pos, ngs, w_t)
                    loss.backward()
                    out[f"{key}/{tag}/pos"] = _np(pos)
                    out[f"{key}/{tag}/neg_score"] = _np(ngs)
                    out[f"{key}/{tag}/loss"] = _np(loss)
                    out[f"{key}/{tag}/grad_ent"] = _np(mdl.entity_embedding.grad)
                    out[f"{key}/{tag}/grad_rel"] = _np(mdl.relation_embedding.grad)
                out[f"{key}/ent"] = _np(model.entity_embedding)
                out[f"{key}/rel"] = _np(model.relation_embedding)
                out[f"{key}/sample"] = sample
                out[f"{key}/neg"] = neg
                out[f"{key}/weight"] = weight
                out[f"{key}/gamma"] = np.float64(gamma)
                # 3-D sample path (base.py:146-151)
                s3 = np.stack([sample[:4], sample[2:6]], axis=0)
                out[f"{key}/sample3d"] = s3
                out[f"{key}/f32/score3d"] = _np(model(torch.from_numpy(s3)))
    np.savez_compressed(os.path.join(HERE, "step_cases.npz"), **out)
    print("step_cases", len(out))


def gen_doctest_pins():
    """The known-answer doctests of the reference, re-executed (not transcribed)."""
    out = {}
    # sampling/negative_sampling.py:62-126
    torch.manual_seed(42)
    entities = {f"e_{i}": i for i in range(4)}
    relations = {f"r_{i}": i for i in range(4)}
    train = [(0, 0, 1), (1, 0, 2), (2, 0, 3), (3, 0, 1)]
    model = models.RotatE(entities=entities, relations=relations, hidden_dim=3, gamma=3)
    dataset = datasets.Dataset(
        train=train, entities=entities, relations=relations, batch_size=2, seed=42, shuffle=False
    )
    ns = sampling.NegativeSampling(
        size=5, train_triples=dataset.train, entities=dataset.entities, relations=dataset.relations, seed=42
    )
    for data in dataset:
        sample = data["sample"]
        break
    neg_tail = ns.generate(sample, mode="tail-batch")
    sc_tail = model(sample, neg_tail, mode="tail-batch")
    neg_head = ns.generate(sample, mode="head-batch")
    sc_head = model(sample, neg_head, mode="head-batch")
    out["ns/ent"] = _np(model.entity_embedding)
    out["ns/rel"] = _np(model.relation_embedding)
    out["ns/sample"] = _np(sample)
    out["ns/neg_tail"] = _np(neg_tail)
    out["ns/score_tail"] = _np(sc_tail)
    out["ns/neg_head"] = _np(neg_head)
    out["ns/score_head"] = _np(sc_head)
    out["ns/train"] = np.array(train, dtype=np.int64)
    # the two pools the sampler drew (same RandomState stream, negative_sampling.py:151,166)
    rs = np.random.RandomState(42)
    out["ns/pool0"] = rs.randint(4, size=10)
    out["ns/pool1"] = rs.randint(4, size=10)
    # the doctest's printed values, for a human reading the fixture
    out["ns/doc_score_tail"] = np.array(
        [[-2.7508, -0.8767, -3.1058, -2.7508, -2.7508], [-2.7456, -0.8674, -2.7456, -0.8674, -0.8674]]
    )
    out["ns/doc_score_head"] = np.array(
        [[-0.3654, -0.3654, -0.3654, -0.3654, -0.3654], [-1.8212, -1.8212, -1.8212, -1.8212, -1.2505]]
    )

    # utils/predict.py:76-95 — TransE on Umls, three scores pinned
    torch.manual_seed(42)
    umls = datasets.Umls(batch_size=2)
    tm = models.TransE(entities=umls.entities, relations=umls.relations, hidden_dim=3, gamma=6)
    q = torch.tensor(umls.test[:3])
    out["predict/ent"] = _np(tm.entity_embedding)
    out["predict/rel"] = _np(tm.relation_embedding)
    out["predict/sample"] = _np(q)
    out["predict/score"] = _np(tm(q)).reshape(-1)
    out["predict/doc_score"] = np.array([-2.4270, -2.1356, -2.4053])
    np.savez_compressed(os.path.join(HERE, "doctest_pins.npz"), **out)
    print("doctest_pins", len(out))


def _toy_graph(rng, N, R, T):
    seen = set()
    while len(seen) < T:
        seen.add((int(rng.randint(N)), int(rng.randint(R)), int(rng.randint(N))))
    return sorted(seen)


def gen_sampler_and_weights():
    """Reference sampler draws, sub-sampling weights and true-set dictionaries on a toy graph."""
    out = {}
    rng = np.random.RandomState(11)
    N, R = 60, 4
    triples = _toy_graph(rng, N, R, 400)
    entities = {f"e{i}": i for i in range(N)}
    relations = {f"r{i}": i for i in range(R)}
    out["triples"] = np.array(triples, dtype=np.int64)
    out["N"], out["R"] = np.int64(N), np.int64(R)
    # weights (datasets/base.py:102-121)
    td = ref_base.TrainDataset(triples=triples, entities=entities, relations=relations, mode="tail-batch", seed=42)
    out["weights"] = np.array([float(td.weights[i]) for i in range(len(triples))], dtype=np.float32)
    # sampler
    size = 16
    ns = sampling.NegativeSampling(size=size, train_triples=triples, entities=entities, relations=relations, seed=42)
    rs = np.random.RandomState(42)
    th_keys = sorted(ns.true_head)
    out["true_head_keys"] = np.array(th_keys, dtype=np.int64)
    out["true_head_sizes"] = np.array([len(ns.true_head[k]) for k in th_keys], dtype=np.int64)
    out["true_head_members"] = np.concatenate([np.sort(ns.true_head[k]) for k in th_keys]).astype(np.int64)
    tt_keys = sorted(ns.true_tail)
    out["true_tail_keys"] = np.array(tt_keys, dtype=np.int64)
    out["true_tail_sizes"] = np.array([len(ns.true_tail[k]) for k in tt_keys], dtype=np.int64)
    out["true_tail_members"] = np.concatenate([np.sort(ns.true_tail[k]) for k in tt_keys]).astype(np.int64)
    for step in range(6):
        mode = "head-batch" if step % 2 == 0 else "tail-batch"
        idx = rng.randint(len(triples), size=8)
        sample = torch.tensor([triples[i] for i in idx])
        negs = ns.generate(sample, mode)
        out[f"gen{step}/sample"] = _np(sample)
        out[f"gen{step}/mode"] = np.array(mode)
        out[f"gen{step}/pool"] = rs.randint(N, size=2 * size)
        out[f"gen{step}/neg"] = _np(negs)
    np.savez_compressed(os.path.join(HERE, "sampler_cases.npz"), **out)
    print("sampler_cases", len(out))


def gen_eval_cases():
    """TestDataset candidate/bias lists, per-query ranks and Evaluation.eval metrics."""
    out = {}
    rng = np.random.RandomState(5)
    N, R = 50, 3
    triples = _toy_graph(rng, N, R, 300)
    train, valid, test = triples[:240], triples[240:270], triples[270:]
    entities = {f"e{i}": i for i in range(N)}
    relations = {f"r{i}": i for i in range(R)}
    out["train"] = np.array(train, dtype=np.int64)
    out["valid"] = np.array(valid, dtype=np.int64)
    out["test"] = np.array(test, dtype=np.int64)
    true_triples = train + valid + test
    for mode in ("head-batch", "tail-batch"):
        td = ref_base.TestDataset(
            triples=test, true_triples=true_triples, entities=entities, relations=relations, mode=mode
        )
        cands, biases = [], []
        for i in range(len(test)):
            _, c, b, _ = td[i]
            cands.append(_np(c))
            biases.append(_np(b))
        out[f"{mode}/cand"] = np.stack(cands)
        out[f"{mode}/bias"] = np.stack(biases)
    for name, gamma in MODEL_GAMMA.items():
        torch.manual_seed(77)
        D = 8
        model = getattr(models, name)(hidden_dim=D, entities=entities, relations=relations, gamma=gamma)
        with torch.no_grad():
            model.entity_embedding.mul_(4.0)
            model.relation_embedding.mul_(4.0)
        ev = evaluation.Evaluation(
            entities=entities, relations=relations, batch_size=4, true_triples=true_triples, num_workers=0
        )
        out[f"{name}/ent"] = _np(model.entity_embedding)
        out[f"{name}/rel"] = _np(model.relation_embedding)
        out[f"{name}/gamma"] = np.float64(gamma)
        metrics = ev.eval(model=model, dataset=test)
        out[f"{name}/metrics"] = np.array([metrics[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")])
        # per-query ranks, same arithmetic as evaluation.py:237-263
        with torch.no_grad():
            for mode in ("head-batch", "tail-batch"):
                s = torch.tensor(test)
                c = torch.from_numpy(out[f"{mode}/cand"])
                b = torch.from_numpy(out[f"{mode}/bias"])
                sc = model(s, c, mode) + b
                order = torch.argsort(sc, dim=1, descending=True)
                pos = s[:, 0] if mode == "head-batch" else s[:, 2]
                ranks = [(order[i] == pos[i]).nonzero().item() + 1 for i in range(len(test))]
                out[f"{name}/{mode}/ranks"] = np.array(ranks, dtype=np.int64)
                out[f"{name}/{mode}/scores"] = _np(sc)
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)
    print("eval_cases", len(out))


def _typed_graph():
    """40 entities, 4 relations with the four mapping categories (r0 one-to-one, r1 one-to-many,
    r2 many-to-one, r3 many-to-many) and a few (head, tail) pairs connected by several relations, so
    that types_relations / detail_eval / TestDatasetRelation's bias all have something to do."""
    rng = np.random.RandomState(12)
    t = set()
    perm = rng.permutation(40)
    for i in range(0, 36, 2):  # r0: disjoint pairs -> 1 head per tail, 1 tail per head
        t.add((int(perm[i]), 0, int(perm[i + 1])))
    for h in (0, 1, 2, 3):  # r1: 4 heads x 6 tails each, tails disjoint -> 1_M
        for k in range(6):
            t.add((h, 1, 4 + 6 * h + k))
    for tl in (30, 31, 32):  # r2: many heads per tail, one tail per head -> M_1
        for k in range(7):
            t.add((7 * (tl - 30) + k, 2, tl))
    for _ in range(70):  # r3: many-to-many
        t.add((int(rng.randint(12)), 3, int(rng.randint(12, 24))))
    t = sorted(t)
    extra = [(h, 3, tl) for h, r, tl in t if r == 1][:6] + [(h, 1, tl) for h, r, tl in t if r == 2][:3]
    t = sorted(set(t) | set(extra))
    order = rng.permutation(len(t))
    t = [t[i] for i in order]
    n_test = 24
    return t[n_test + 10:], t[n_test:n_test + 10], t[:n_test]


def gen_next_rows():
    """TestDatasetRelation items, eval_relations, types_relations / detail_eval,
    utils.TopK, utils.make_prediction, pRotatE forward/backward (incl. the trainable modulus),
    KlDivergence value and gradient."""
    from mkb import utils
    from mkb.models import pRotatE

    out = {}
    N, R = 40, 4
    train, valid, test = _typed_graph()
    entities = {f"e{i}": i for i in range(N)}
    relations = {f"r{i}": i for i in range(R)}
    true_triples = train + valid + test
    out["train"], out["valid"], out["test"] = (np.array(x, dtype=np.int64) for x in (train, valid, test))
    tdr = ref_base.TestDatasetRelation(triples=test, true_triples=true_triples, entities=entities, relations=relations)
    items = [tdr[i] for i in range(len(test))]
    out["rel/cand"] = np.stack([_np(c) for _, c, _, _ in items])
    out["rel/bias"] = np.stack([_np(b) for _, _, b, _ in items])
    for name, gamma in MODEL_GAMMA.items():
        torch.manual_seed(77)
        model = getattr(models, name)(hidden_dim=8, entities=entities, relations=relations, gamma=gamma)
        with torch.no_grad():
            model.entity_embedding.mul_(4.0)
            model.relation_embedding.mul_(4.0)
        out[f"{name}/ent"], out[f"{name}/rel"] = _np(model.entity_embedding), _np(model.relation_embedding)
        out[f"{name}/gamma"] = np.float64(gamma)
        ev = evaluation.Evaluation(entities=entities, relations=relations, batch_size=4, true_triples=true_triples,
                                   num_workers=0)
        m = ev.eval_relations(model=model, dataset=test)
        out[f"{name}/rel_metrics"] = np.array([m[f"{k}_relations"] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")])
        types = ev.types_relations(model=model, dataset=test)
        out[f"{name}/types"] = np.array([types[f"r{i}"] for i in range(R)])
        frame = ev.detail_eval(model=model, dataset=test)
        out[f"{name}/detail"] = frame.to_numpy(dtype=np.float64)
        out[f"{name}/detail_cols"] = np.array(["|".join(c) for c in frame.columns])
        out[f"{name}/detail_index"] = np.array(list(frame.index))
        topk = utils.TopK(entities=entities, relations=relations)
        out[f"{name}/top_heads"] = np.array([entities[e] for e in topk.top_heads(k=7, model=model, relation="r1", tail="e3")])
        out[f"{name}/top_tails"] = np.array([entities[e] for e in topk.top_tails(k=7, model=model, head="e5", relation=2)])
        out[f"{name}/top_relations"] = np.array([relations[r] for r in topk.top_relations(k=2, model=model, head=4, tail="e9")])
        out[f"{name}/prediction"] = _np(utils.make_prediction(model=model, dataset=test[:9], batch_size=4, num_workers=0,
                                                              device="cpu"))
        # triplet classification (evaluation/classif.py): true test triples vs the same with a shifted tail
        X = list(test) + [(h, r, (t + 7) % N) for h, r, t in test]
        y = [1] * len(test) + [-1] * len(test)
        out["clf/X"], out["clf/y"] = np.array(X, dtype=np.int64), np.array(y, dtype=np.int64)
        thr = evaluation.find_threshold(model=model, X=X, y=y, batch_size=8, num_workers=0, device="cpu")
        out[f"{name}/threshold"] = np.float64(thr)
        out[f"{name}/accuracy"] = np.float64(evaluation.accuracy(model=model, X=X, y=y, threshold=thr, batch_size=8,
                                                                 num_workers=0, device="cpu"))
    # pRotatE: forward / loss / autograd gradients (fp32 and fp64), both modes, vector and scalar dims
    for D in (8, 5):
        for mode in ("tail-batch", "head-batch"):
            torch.manual_seed(11 + D)
            model = pRotatE(hidden_dim=D, entities=entities, relations=relations, gamma=9.0)
            with torch.no_grad():
                model.entity_embedding.mul_(3.0)
                model.relation_embedding.mul_(3.0)
            r2 = np.random.RandomState(D)
            sample = torch.tensor(np.stack([r2.randint(N, size=6), r2.randint(R, size=6), r2.randint(N, size=6)], 1))
            neg = torch.tensor(r2.randint(N, size=(6, 7)))
            weight = torch.tensor(r2.uniform(0.1, 0.5, 6).astype(np.float32))
            key = f"pRotatE_D{D}_{mode}"
            out[f"{key}/ent"], out[f"{key}/rel"] = _np(model.entity_embedding).copy(), _np(model.relation_embedding).copy()
            out[f"{key}/modulus"] = _np(model.modulus).copy()
            out[f"{key}/sample"], out[f"{key}/neg"], out[f"{key}/weight"] = _np(sample), _np(neg), _np(weight)
            for tag, mdl in (("f32", model), ("f64", __import__("copy").deepcopy(model).double())):
                w = weight.double() if tag == "f64" else weight
                mdl.zero_grad()
                pos = mdl(sample)
                ns = mdl(sample, neg, mode)
                loss = losses.Adversarial(alpha=0.5)(pos, ns, w)
                loss.backward()
                out[f"{key}/{tag}/pos"], out[f"{key}/{tag}/neg_score"] = _np(pos), _np(ns)
                out[f"{key}/{tag}/loss"] = np.float64(loss.item())
                out[f"{key}/{tag}/grad_ent"] = _np(mdl.entity_embedding.grad)
                out[f"{key}/{tag}/grad_rel"] = _np(mdl.relation_embedding.grad)
                out[f"{key}/{tag}/grad_modulus"] = _np(mdl.modulus.grad)
                out[f"{key}/{tag}/score3d"] = _np(mdl(torch.stack([sample[:4], sample[2:6]])))
    # KlDivergence (losses/kl_divergence.py:22-29): value and gradient w.r.t. the student scores
    r3 = np.random.RandomState(9)
    for T in (1, 3):
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            student = torch.tensor(r3.normal(0, 2, size=(5, 11)), dtype=dt, requires_grad=True)
            teacher = torch.tensor(r3.normal(0, 2, size=(5, 11)), dtype=dt)
            loss = losses.KlDivergence()(student, teacher, T=T)
            loss.backward()
            out[f"kl_T{T}/{tag}/student"], out[f"kl_T{T}/{tag}/teacher"] = _np(student), _np(teacher)
            out[f"kl_T{T}/{tag}/loss"] = np.float64(loss.item())
            out[f"kl_T{T}/{tag}/grad"] = _np(student.grad)
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
    print("next_rows", len(out))



def gen_eval_doctest():
    """evaluation/evaluation.py:39-116 — train RotatE(dim 3) for 5 epochs on a 4-triple toy graph
    with the doctest's own loop (note: it never calls zero_grad, so .grad accumulates), then pin
    MRR/MR/HITS.  Stores the per-step inputs so the replacement can replay the identical
    sequence, plus the trained tables and metrics."""
    out = {}
    torch.manual_seed(42)
    train = [(0, 0, 1), (0, 1, 1), (2, 0, 3), (2, 1, 3)]
    valid = [(0, 0, 1), (2, 1, 3)]
    test = [(0, 0, 1), (2, 1, 3)]
    entities = {"e0": 0, "e1": 1, "e2": 2, "e3": 3}
    relations = {"r0": 0, "r1": 1}
    dataset = datasets.Dataset(
        train=train, valid=valid, test=test, entities=entities, relations=relations,
        batch_size=2, seed=42, shuffle=False,
    )
    negative_sampling = sampling.NegativeSampling(
        size=2, train_triples=dataset.train, entities=dataset.entities, relations=dataset.relations, seed=42
    )
    model = models.RotatE(hidden_dim=3, entities=dataset.entities, relations=dataset.relations, gamma=1)
    out["ent0"] = _np(model.entity_embedding).copy()
    out["rel0"] = _np(model.relation_embedding).copy()
    optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=0.5)
    loss = losses.Adversarial(alpha=0.5)
    steps = 0
    for _ in range(5):
        for data in dataset:
            sample, weight, mode = data["sample"], data["weight"], data["mode"]
            positive_score = model(sample)
            negative_sample = negative_sampling.generate(sample=sample, mode=mode)
            negative_score = model(sample, negative_sample, mode)
            err = loss(positive_score, negative_score, weight)
            err.backward()
            optimizer.step()
            out[f"step{steps}/sample"] = _np(sample)
            out[f"step{steps}/weight"] = _np(weight)
            out[f"step{steps}/neg"] = _np(negative_sample)
            out[f"step{steps}/mode"] = np.array(mode)
            out[f"step{steps}/loss"] = _np(err)
            steps += 1
    out["n_steps"] = np.int64(steps)
    out["ent_final"] = _np(model.entity_embedding)
    out["rel_final"] = _np(model.relation_embedding)
    out["train"] = np.array(train, dtype=np.int64)
    out["test"] = np.array(test, dtype=np.int64)
    model = model.eval()
    validation = evaluation.Evaluation(
        true_triples=train + valid + test, entities=entities, relations=relations, batch_size=2
    )
    m = validation.eval(model=model, dataset=test)
    out["metrics"] = np.array([m[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")])
    out["doc_metrics"] = np.array([0.5417, 2.25, 0.25, 1.0, 1.0])
    np.savez_compressed(os.path.join(HERE, "eval_doctest.npz"), **out)
    print("eval_doctest", len(out), m)


def gen_loader_order():
    """Batch order of the reference's Dataset(shuffle=True) under a fixed torch seed, for both
    DataLoader regimes (worker process / in-process), two epochs each."""
    out = {}
    rng = np.random.RandomState(0)
    tri = sorted({(int(rng.randint(50)), int(rng.randint(3)), int(rng.randint(50))) for _ in range(300)})
    ents, rels = {i: i for i in range(50)}, {i: i for i in range(3)}
    out["triples"] = np.array(tri, dtype=np.int64)
    for nw in (1, 0):
        torch.manual_seed(123)
        ds = datasets.Dataset(train=tri, entities=ents, relations=rels, batch_size=32, shuffle=True, seed=42,
                              num_workers=nw)
        for epoch in range(2):
            batches = list(ds)
            out[f"nw{nw}/epoch{epoch}/n"] = np.int64(len(batches))
            for i, d in enumerate(batches):
                out[f"nw{nw}/epoch{epoch}/b{i}"] = _np(d["sample"])
                out[f"nw{nw}/epoch{epoch}/m{i}"] = np.array(d["mode"])
    np.savez_compressed(os.path.join(HERE, "loader_order.npz"), **out)
    print("loader_order", len(out))


def gen_distill_rows():
    """TopKSampling.get of the reference (distillation/top_k_sampling.py:564-660) on two small KGs whose
    label sets overlap only partly, for a RotatE and a TransE teacher; stored: tables, label maps, outputs."""
    from mkb import distillation

    out = {}
    rng = np.random.RandomState(9)
    Nt, Ns, Rt, Rs, D = 60, 50, 7, 5, 8
    labels_t = [f"e{i}" for i in rng.permutation(80)[:Nt]]
    labels_s = [f"e{i}" for i in rng.permutation(80)[:Ns]]
    rl_t = [f"r{i}" for i in rng.permutation(9)[:Rt]]
    rl_s = [f"r{i}" for i in rng.permutation(9)[:Rs]]
    ent_t = {e: i for i, e in enumerate(labels_t)}
    ent_s = {e: i for i, e in enumerate(labels_s)}
    rel_t = {r: i for i, r in enumerate(rl_t)}
    rel_s = {r: i for i, r in enumerate(rl_s)}
    out["labels_t"], out["labels_s"] = np.array(labels_t), np.array(labels_s)
    out["rl_t"], out["rl_s"] = np.array(rl_t), np.array(rl_s)
    sample = torch.tensor(np.stack([rng.randint(Nt, size=6), rng.randint(Rt, size=6), rng.randint(Nt, size=6)], 1))
    out["sample"] = _np(sample)
    for name in ("RotatE", "TransE", "ComplEx"):
        torch.manual_seed(3)
        teacher = getattr(models, name)(hidden_dim=D, entities=ent_t, relations=rel_t, gamma=6)
        with torch.no_grad():
            teacher.entity_embedding.mul_(3.0)
        out[f"{name}/ent"], out[f"{name}/rel"] = _np(teacher.entity_embedding), _np(teacher.relation_embedding)
        for tag, (ke, kr, ne, nr) in {"a": (4, 2, 2, 1), "b": (7, 3, 0, 0)}.items():
            smp = distillation.TopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                            student_relations=rel_s, batch_size_entity=ke, batch_size_relation=kr,
                                            n_random_entities=ne, n_random_relations=nr, seed=42)
            for call in range(2):  # two calls: the RNG stream continues
                res = smp.get(sample=sample, teacher=teacher)
                for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), res):
                    out[f"{name}/{tag}/{call}/{k}"] = _np(t)
            assert teacher.training
    # Distillation.distill (distillation/distillation.py:440-677): loss value + the student's autograd gradients
    # for a supervised (UniformSampling) and an unsupervised (TopKSampling) sampler, teacher / student KGs that
    # share only part of their labels (so some positives take part in one, two or none of the three terms)
    sample2 = torch.tensor(np.stack([rng.randint(Nt, size=12), rng.randint(Rt, size=12), rng.randint(Nt, size=12)], 1))
    out["sample2"] = _np(sample2)
    for name in ("RotatE", "TransE", "ComplEx"):
        torch.manual_seed(3)
        teacher = getattr(models, name)(hidden_dim=D, entities=ent_t, relations=rel_t, gamma=6)
        with torch.no_grad():
            teacher.entity_embedding.mul_(3.0)
        torch.manual_seed(5)
        student = getattr(models, name)(hidden_dim=D, entities=ent_s, relations=rel_s, gamma=6)
        with torch.no_grad():
            student.entity_embedding.mul_(3.0)
        out[f"{name}/s_ent"], out[f"{name}/s_rel"] = _np(student.entity_embedding), _np(student.relation_embedding)
        for kind in ("uniform", "topk"):
            if kind == "uniform":
                smp = distillation.UniformSampling(batch_size_entity=5, batch_size_relation=3, seed=42)
            else:
                smp = distillation.TopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                                student_relations=rel_s, batch_size_entity=4, batch_size_relation=2,
                                                n_random_entities=2, n_random_relations=1, seed=42)
            proc = distillation.Distillation(teacher_entities=ent_t, student_entities=ent_s, teacher_relations=rel_t,
                                             student_relations=rel_s, sampling=smp)
            av = [proc.available(*map(int, row)) for row in sample2]
            out[f"{name}/{kind}/avail"] = np.array([[a["head"], a["relation"], a["tail"]] for a in av])
            for call in range(2):
                student.zero_grad()
                loss = proc.distill(teacher=teacher, student=student, sample=sample2)
                loss.backward()
                out[f"{name}/{kind}/{call}/loss"] = _np(loss)
                out[f"{name}/{kind}/{call}/g_ent"] = _np(student.entity_embedding.grad)
                out[f"{name}/{kind}/{call}/g_rel"] = _np(student.relation_embedding.grad)
    # FastTopKSampling (:10-318): pre-computed over the teacher's training set, looked up per batch
    train = sorted({(int(rng.randint(Nt)), int(rng.randint(Rt)), int(rng.randint(Nt))) for _ in range(40)})
    out["fast/train"] = np.array(train)
    for name in ("RotatE", "ComplEx"):
        torch.manual_seed(3)
        teacher = getattr(models, name)(hidden_dim=D, entities=ent_t, relations=rel_t, gamma=6)
        with torch.no_grad():
            teacher.entity_embedding.mul_(3.0)
        ds = datasets.Dataset(train=train, entities=ent_t, relations=rel_t, batch_size=7, shuffle=False, seed=42)
        smp = distillation.FastTopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                            student_relations=rel_s, batch_size_entity=4, batch_size_relation=2,
                                            n_random_entities=2, n_random_relations=1, seed=42, teacher=teacher,
                                            dataset_teacher=ds)
        q = torch.tensor(train[3:29:5])
        out["fast/query"] = _np(q)
        for call in range(2):
            res = smp.get(sample=q)
            for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), res):
                out[f"fast/{name}/{call}/{k}"] = _np(t)
    # KdmkbModel.forward (kdmkb_model.py:286-360): two KBs with partly shared labels trained together for a few
    # steps; per-step losses of both models and the final tables
    train_s = sorted({(int(rng.randint(Ns)), int(rng.randint(Rs)), int(rng.randint(Ns))) for _ in range(36)})
    out["kd/train_s"] = np.array(train_s)
    kd_models = {"a": ("RotatE", ent_t, rel_t, train, 3), "b": ("ComplEx", ent_s, rel_s, train_s, 5)}
    ms, dss = collections.OrderedDict(), collections.OrderedDict()
    for key, (name, ents_, rels_, tr, sd) in kd_models.items():
        torch.manual_seed(sd)
        ms[key] = getattr(models, name)(hidden_dim=D, entities=ents_, relations=rels_, gamma=6)
        with torch.no_grad():
            ms[key].entity_embedding.mul_(3.0)
        # copies: _np() aliases the parameter's storage, which the optimizer then updates in place
        out[f"kd/{key}/ent0"] = _np(ms[key].entity_embedding).copy()
        out[f"kd/{key}/rel0"] = _np(ms[key].relation_embedding).copy()
        dss[key] = datasets.Dataset(train=tr, valid=tr[:6], test=tr[6:12], entities=ents_, relations=rels_, batch_size=6,
                                    shuffle=False, seed=42)
    per = lambda v: {"a": v, "b": v}
    kd = distillation.KdmkbModel(models=ms, datasets=dss, lr=per(0.01), alpha_kl=per(0.4), alpha_adv=per(0.5),
                                 negative_sampling_size=per(5), batch_size_entity=per(4), batch_size_relation=per(2),
                                 n_random_entities=per(2), n_random_relations=per(1), update_distillation_every=1000,
                                 device="cpu", seed=42, warm_step=0)
    losses_ = []
    for step in range(4):
        m = kd.forward(dss, ms, weight_kl={"a": 0.4, "b": 0.4})
        losses_.append([m["a"].get(), m["b"].get()])
    out["kd/rolling_loss"] = np.array(losses_)
    for key in ms:
        out[f"kd/{key}/ent1"], out[f"kd/{key}/rel1"] = _np(ms[key].entity_embedding), _np(ms[key].relation_embedding)
    np.savez_compressed(os.path.join(HERE, "distill_rows.npz"), **out)
    print("distill_rows", len(out))


def gen_distill_doctests():
    """The reference's own known answers for the distillation samplers / loss (doctests at
    distillation/top_k_sampling.py:352-413 and distillation/distillation.py:476-522), re-executed here: the inputs
    (label maps in dict order, seeded tables, the sample) are stored so the CUDA path can be held to the numbers
    PRINTED in those docstrings; the asserts below confirm the reference still reproduces them under this torch."""
    from mkb import distillation

    out = {}

    def maps(prefix, d):
        out[f"{prefix}/labels"] = np.array(list(d.keys()))
        out[f"{prefix}/ids"] = np.array(list(d.values()), dtype=np.int64)

    torch.manual_seed(42)
    dt = datasets.CountriesS1(batch_size=2, seed=42, shuffle=False)
    dsd = datasets.CountriesS2(batch_size=2, seed=42, shuffle=False)
    teacher = models.RotatE(entities=dt.entities, relations=dt.relations, gamma=3, hidden_dim=4)
    maps("topk/ent_t", dt.entities), maps("topk/ent_s", dsd.entities)
    maps("topk/rel_t", dt.relations), maps("topk/rel_s", dsd.relations)
    out["topk/ent"], out["topk/rel"] = _np(teacher.entity_embedding).copy(), _np(teacher.relation_embedding).copy()
    smp = distillation.TopKSampling(teacher_relations=dt.relations, teacher_entities=dt.entities,
                                    student_entities=dsd.entities, student_relations=dsd.relations, batch_size_entity=4,
                                    batch_size_relation=1, n_random_entities=1, n_random_relations=0, seed=42)
    sample = next(iter(dt))["sample"]
    assert sample.tolist() == [[0, 0, 266], [1, 1, 56]]
    out["topk/sample"] = _np(sample)
    res = smp.get(sample=sample, teacher=teacher)
    assert res[0].tolist() == [[197, 50, 75, 176, 30], [10, 240, 251, 3, 30]]  # top_k_sampling.py:389-391
    assert res[5].tolist() == [[269, 198, 270, 256, 28], [111, 149, 201, 234, 28]]  # :411-413
    for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), res):
        out[f"topk/{k}"] = _np(t)
    # FastTopKSampling doctest (:24-81): same teacher and sample, tables pre-computed over CountriesS1's training set
    out["topk/train"] = np.array(dt.train, dtype=np.int64)
    fast = distillation.FastTopKSampling(teacher_relations=dt.relations, teacher_entities=dt.entities,
                                         student_entities=dsd.entities, student_relations=dsd.relations,
                                         batch_size_entity=4, batch_size_relation=1, n_random_entities=1,
                                         n_random_relations=0, seed=42, teacher=teacher, dataset_teacher=dt)
    res = fast.get(sample=sample, teacher=teacher)
    assert res[0].tolist() == [[197, 50, 75, 176, 30], [10, 240, 251, 3, 30]]  # :53-55
    assert res[5].tolist() == [[269, 198, 270, 256, 28], [111, 149, 201, 234, 28]]  # :77-79

    torch.manual_seed(42)
    ds = datasets.Umls(batch_size=3, shuffle=False, seed=42)
    teacher = models.RotatE(hidden_dim=3, entities=ds.entities, relations=ds.relations, gamma=6)
    student = models.RotatE(hidden_dim=3, entities=ds.entities, relations=ds.relations, gamma=6)
    maps("umls/ent", ds.entities), maps("umls/rel", ds.relations)
    out["umls/t_ent"], out["umls/t_rel"] = _np(teacher.entity_embedding).copy(), _np(teacher.relation_embedding).copy()
    out["umls/s_ent"], out["umls/s_rel"] = _np(student.entity_embedding).copy(), _np(student.relation_embedding).copy()
    proc = distillation.Distillation(teacher_entities=ds.entities, student_entities=ds.entities,
                                     teacher_relations=ds.relations, student_relations=ds.relations,
                                     sampling=distillation.UniformSampling(batch_size_entity=3, batch_size_relation=3,
                                                                           seed=42))
    sample = next(iter(ds))["sample"]
    out["umls/sample"] = _np(sample)
    loss = proc.distill(teacher=teacher, student=student, sample=sample)
    assert round(loss.item(), 4) == 1.3066  # distillation.py:500-501
    out["umls/loss"] = _np(loss)
    np.savez_compressed(os.path.join(HERE, "distill_doctests.npz"), **out)
    print("distill_doctests", len(out))


CFG5_MARGINS = (1e-6, 1e-5, 1e-4)  # relative to max(|s_pos|, mean|s|), see oracle rank_all


def gen_cfg5_cases(n_queries=64, n_queries_rotate=32):
    """BASELINE config 5 AT SIZE: the reference's own filtered ranks on the real Wn18rr (N = 40 943, 11
    relations, filter = train + valid + test) at D = 1000, from its unmodified Evaluation.compute_score
    (evaluation/evaluation.py:217-279 over datasets/base.py:196-241's candidate / filter_bias lists).
    Tables are rebuilt from a seed on both sides (tests/golden/cfg5_tables.py).  Per model: the queries,
    the reference's fp32 ranks for both modes, the fp64 oracle's ranks and, per query, how many unfiltered
    candidates lie within 1e-6 / 1e-5 / 1e-4 (relative) of the positive's score (the near-ties an fp32 implementation with a
    different summation order may legitimately flip)."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), ".."))
    sys.path.insert(0, HERE)
    import cfg5_tables
    from oracle import kge_oracle as ko
    from river import stats as rstats

    class Rec(rstats.Mean):
        def __init__(self):
            super().__init__()
            self.seen = []

        def update(self, x):
            self.seen.append(x)
            return super().update(x)

    ds = datasets.Wn18rr(batch_size=1, shuffle=False, pre_compute=False, seed=42)
    N, R, D = ds.n_entity, ds.n_relation, 1000
    train = np.array(ds.train, dtype=np.int64)
    true_triples = ds.train + ds.valid + ds.test
    all_true = np.array(true_triples, dtype=np.int64)
    hc, tc = ko.build_filter_csr(all_true, N, "head"), ko.build_filter_csr(all_true, N, "tail")
    out = {"n_entity": np.int64(N), "n_relation": np.int64(R), "hidden_dim": np.int64(D),
           "train": train.astype(np.int32), "valid": np.array(ds.valid, dtype=np.int32),
           "test": np.array(ds.test, dtype=np.int32)}
    pick = np.random.RandomState(5).permutation(len(ds.test))
    for name, gamma in MODEL_GAMMA.items():
        nq = n_queries_rotate if name == "RotatE" else n_queries
        queries = [ds.test[i] for i in pick[:nq]]
        ent, rel = cfg5_tables.make_tables(name, train, N, R, D, gamma)
        model = getattr(models, name)(hidden_dim=D, entities=ds.entities, relations=ds.relations, gamma=gamma)
        with torch.no_grad():
            model.entity_embedding.copy_(torch.from_numpy(ent))
            model.relation_embedding.copy_(torch.from_numpy(rel))
        ev = evaluation.Evaluation(entities=ds.entities, relations=ds.relations, batch_size=2,
                                   true_triples=true_triples, num_workers=0)
        metrics = collections.OrderedDict({m: Rec() for m in ["MRR", "MR", "HITS@1", "HITS@3", "HITS@10"]})
        import time as _t
        t0 = _t.time()
        with torch.no_grad():
            for stream in ev.get_entity_stream(queries):  # head-batch stream, then tail-batch stream
                metrics = ev.compute_score(model=model, test_set=stream, metrics=metrics, device="cpu")
        ranks = np.array(metrics["MR"].seen, dtype=np.int64).reshape(2, nq)  # [head-batch, tail-batch]
        t_ref = _t.time() - t0
        out[f"{name}/queries"] = np.array(queries, dtype=np.int64)
        out[f"{name}/gamma"] = np.float64(gamma)
        out[f"{name}/ref_ranks"] = ranks
        out[f"{name}/ref_metrics"] = np.array([round(m.get(), 4) for m in metrics.values()])
        out[f"{name}/ent_checksum"] = np.float64(ent.astype(np.float64).sum())
        out[f"{name}/rel_checksum"] = np.float64(rel.astype(np.float64).sum())
        for mi, mode in enumerate(("head-batch", "tail-batch")):
            r64, cont = ko.rank_all(name, ent, rel, np.array(queries), mode, hc, tc, gamma=gamma,
                                    rel_margins=CFG5_MARGINS)
            out[f"{name}/{mode}/rank64"] = r64
            out[f"{name}/{mode}/contested"] = cont  # [Q, len(CFG5_MARGINS)]
            print(name, mode, "ref==fp64:", int((ranks[mi] == r64).sum()), "/", nq, "contested>0:",
                  (cont > 0).sum(0).tolist(), "max |ref-fp64|", int(np.abs(ranks[mi] - r64).max()),
                  "median rank", int(np.median(r64)), f"ref eval {t_ref:.0f}s", flush=True)
    out["margins"] = np.array(CFG5_MARGINS)
    np.savez_compressed(os.path.join(HERE, "cfg5_wn18rr.npz"), **out)
    print("cfg5_wn18rr", len(out))


def gen_cfg1_epoch():
    """BASELINE config 1 exactly as worded — "Wn18rr TransE dim=200 batch=256 neg=64 Adversarial loss on CPU
    (reference compose.Pipeline, 1 epoch)" — run through the UNMODIFIED reference: mkb.datasets.Wn18rr,
    models.TransE, sampling.NegativeSampling, losses.Adversarial, torch.optim.Adam, compose.Pipeline(epochs=1).learn
    (compose/pipeline.py:202-244).  Stored: the loss of every step (recorded by wrapping the loss object), the
    trained relation table, 512 sampled rows of the trained entity table + a checksum of the whole table, and the
    seeds — so the CUDA path can replay the epoch (same loader order, same reference-pool negatives) and must land
    on the same trajectory."""
    from mkb import compose

    out = {}
    D, B, K, gamma, lr, seed = 200, 256, 64, 6.0, 5e-5, 42
    torch.manual_seed(seed)
    ds = datasets.Wn18rr(batch_size=B, shuffle=True, seed=seed)
    model = models.TransE(hidden_dim=D, entities=ds.entities, relations=ds.relations, gamma=gamma)
    out["ent0_checksum"] = np.float64(_np(model.entity_embedding).astype(np.float64).sum())
    sampler = sampling.NegativeSampling(size=K, train_triples=ds.train, entities=ds.entities, relations=ds.relations,
                                        seed=seed)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)
    base = losses.Adversarial(alpha=0.5)
    seen = []

    def rec_loss(positive_score, negative_score, weight):
        err = base(positive_score, negative_score, weight)
        seen.append(float(err.detach()))
        return err

    import time as _t
    t0 = _t.time()
    pipe = compose.Pipeline(epochs=1)
    old = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        class _NoEval:  # the reference dereferences `evaluation` unguarded after the last epoch
            def eval(self, model, dataset):
                return {}

            def eval_relations(self, model, dataset):
                return {}

        pipe.learn(model=model, dataset=ds, sampling=sampler, optimizer=opt, loss=rec_loss, evaluation=_NoEval())
    finally:
        sys.stdout = old
    ent, rel = _np(model.entity_embedding), _np(model.relation_embedding)
    rows = np.random.RandomState(7).choice(ent.shape[0], 512, replace=False)
    out.update({"hidden_dim": np.int64(D), "batch": np.int64(B), "neg": np.int64(K), "gamma": np.float64(gamma),
                "lr": np.float64(lr), "seed": np.int64(seed), "losses": np.array(seen, dtype=np.float64),
                "rel_final": rel, "rows": rows.astype(np.int64), "ent_rows_final": ent[rows],
                "ent_checksum": np.float64(ent.astype(np.float64).sum()),
                "ent_abs_checksum": np.float64(np.abs(ent).astype(np.float64).sum()),
                "rolling_loss": np.float64(pipe.metric_loss.get())})
    np.savez_compressed(os.path.join(HERE, "cfg1_epoch.npz"), **out)
    print("cfg1_epoch", len(seen), "steps in", f"{_t.time() - t0:.0f}s; loss", seen[0], "->", seen[-1])


if __name__ == "__main__":
    which = sys.argv[1:] or ["step", "pins", "sampler", "eval", "evaldoc", "loader", "next", "distill", "distilldoc"]
    if "cfg1" in which:  # ~1-2 min of CPU: one reference epoch of config 1
        gen_cfg1_epoch()
        sys.exit(0)
    if "cfg5" in which:  # ~20 min of CPU (the reference ranks 40 943 candidates x D=1000 per query)
        gen_cfg5_cases(*[int(a) for a in which[which.index("cfg5") + 1:][:2]])
        sys.exit(0)
    if "next" in which:
        gen_next_rows()
    if "distill" in which:
        gen_distill_rows()
    if "distilldoc" in which:
        gen_distill_doctests()
    if "loader" in which:
        gen_loader_order()
    if "step" in which:
        gen_step_cases()
    if "pins" in which:
        gen_doctest_pins()
    if "sampler" in which:
        gen_sampler_and_weights()
    if "eval" in which:
        gen_eval_cases()
    if "evaldoc" in which:
        gen_eval_doctest()
