"""Build container only (needs /root/reference): run EVERY `examples/*/problem.py` of the reference, unmodified,
through `multipinn_b200.install()` and compare the fused step (traced residual programs, oracle backend standing in
for the kernels -- there is no GPU here) with the unpatched reference on the same weights and points:

  * per-condition losses of `PINN.calculate_loss` (reference `multipinn/PINN/PINN.py:109-125`),
  * the flat parameter gradient of the mean loss,

and report, per condition, whether the residual was traced, the jet mode it runs in and the engine that would take it.
Examples that need files that are not in the tree (meshes, data) or the graph network are listed as skipped.

    python scripts/example_sweep.py [name ...]  > profiles/r2_example_sweep.txt
"""
import importlib
import importlib.util
import inspect
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mk)
ref = mk.import_reference()
import multipinn_b200  # noqa: E402
from multipinn_b200 import engine  # noqa: E402

ob = importlib.util.spec_from_file_location("ob", os.path.join(ROOT, "tests", "oracle_backend.py"))
obm = importlib.util.module_from_spec(ob)
ob.loader.exec_module(obm)
engine._set_backend_for_tests(obm.OracleBackend())

N_PTS = int(os.environ.get("SWEEP_POINTS", "24"))
HIDDEN = [64, 64, 64]


def engine_supports(eng, order):
    """Host-side query of the real library (no GPU needed): does the tcgen05 engine take this program?"""
    import ctypes
    from multipinn_b200 import _lib
    lib = _lib.load()
    return bool(lib.mpinn_engine_supports(eng.ops.ctypes.data_as(ctypes.c_void_p), int(eng.ops.shape[0]),
                                          int(eng.d), int(order), int(_lib.ENGINE_TCGEN05)))


def find_factory(mod):
    """The function of the module that returns (conditions, input_dim, output_dim) when called with its defaults."""
    cands = [f for n, f in inspect.getmembers(mod, inspect.isfunction) if f.__module__ == mod.__name__]
    cands.sort(key=lambda f: -f.__code__.co_firstlineno)     # the factory is usually the last function
    for f in cands:
        sig = inspect.signature(f)
        if any(p.default is inspect._empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)
               for p in sig.parameters.values()):
            if list(sig.parameters) == ["c"]:     # wave examples: wave speed
                return lambda f=f: f(1.0)
            continue
        return f
    return None


def fresh_import(modname):
    for k in [k for k in sys.modules if k == modname or k.startswith("examples.")]:
        del sys.modules[k]
    return importlib.import_module(modname)


def build(modname, state=None, pts=None):
    mod = fresh_import(modname)
    fac = find_factory(mod)
    if fac is None:
        raise RuntimeError("no factory with defaults")
    out = fac()
    if not (isinstance(out, tuple) and len(out) >= 3 and isinstance(out[0], (list, tuple))):
        raise RuntimeError(f"factory returned {type(out)}")
    conds, din, dout = out[:3]          # (marching / gradual-learning examples append their schedule objects)
    if not (isinstance(din, int) and isinstance(dout, int)):
        ints = [v for v in out[1:] if isinstance(v, int)]
        din, dout = ints[:2]
    import multipinn
    torch.manual_seed(0)
    model = multipinn.neural_network.FNN(input_dim=din, output_dim=dout, hidden_layers=HIDDEN)
    if state is not None:
        model.load_state_dict(state)
    if pts is None:
        rng_state = np.random.get_state()
        np.random.seed(0)
        pts = []
        for c in conds:
            p = np.asarray(c.geometry.random_points(N_PTS, "pseudo"), dtype=np.float32)
            pts.append(p)
        np.random.set_state(rng_state)
    for c, p in zip(conds, pts):
        c.generator = mk._FixedGenerator(torch.tensor(p).requires_grad_())
    return mod, conds, model, pts


def one(name):
    modname = f"examples.{name}.problem"
    # ---- unpatched reference
    multipinn_b200.uninstall()
    _, conds_r, model_r, pts = build(modname)
    import multipinn
    pinn_r = multipinn.PINN(model=model_r, conditions=conds_r)
    pinn_r.update_data()
    for c in conds_r:
        c.batch_size = len(c.points)
    pinn_r.select_batch(0)
    losses_r = pinn_r.calculate_loss()
    losses_r = torch.stack(list(losses_r)) if not torch.is_tensor(losses_r) else losses_r
    losses_r.mean().backward()
    grad_r = torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel())
                        for p in model_r.parameters()]).numpy()
    state = {k: v.detach().clone() for k, v in model_r.state_dict().items()}
    # ---- patched: the same unmodified file on the fused step
    multipinn_b200.install()
    _, conds, model, _ = build(modname, state, pts)
    import multipinn
    pinn = multipinn.PINN(model=model, conditions=conds)
    pinn.update_data()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    tr = multipinn.TrainerOneBatch(pinn, opt, sched, num_epochs=1, update_grid_every=10 ** 9, calc_loss="mean")
    notes = []
    for c in conds:
        prog = c.program_for(model)
        if prog is None:
            notes.append(f"EAGER ({c.trace_note})")
        else:
            mode = {0: "value", 1: "first", 2: "second", 3: "laplace", 4: "hessian"}
            tc = lambda b: "tcgen05" if engine_supports(model.jet_engine, b.order) else "ffma"   # noqa: E731
            notes.append(f"traced, {len(prog.instrs)} instr: " + " + ".join(
                f"model call [{mode[b.order]}{'' if b.lap is None else ' c=' + str(tuple(round(c, 4) for c in b.lap))}"
                f"{'' if b.points is None else ', derived points'}] -> {tc(b)}"
                for b in prog.blocks) if prog.blocks else "traced (no model call)")
    fs = tr.fused_step()
    if fs is not None:
        for c in conds:
            c.batch_size = len(c.points)
        pinn.select_batch(0)
        fs.compute()
        losses = fs.losses()[1].detach().cpu().numpy().astype(np.float64)
        g = model.jet_engine.flat_grad.detach().cpu().numpy().copy()
        path = "fused"
    else:
        for c in conds:
            c.batch_size = len(c.points)
        pinn.select_batch(0)
        ls = pinn.calculate_loss()
        ls = torch.stack(list(ls)) if not torch.is_tensor(ls) else ls
        model.zero_grad()
        ls.mean().backward()
        losses = ls.detach().numpy().astype(np.float64)
        g = model.jet_engine.flat_grad.detach().numpy().copy()
        path = "eager"
    lr = losses_r.detach().numpy().astype(np.float64)
    el = np.abs(losses - lr).max() / max(np.abs(lr).max(), 1e-30)
    # flat layouts: the mirror packs parameters in state-dict order as the reference iterates them
    eg = float("nan")
    if g is not None and g.shape == grad_r.shape:
        eg = np.abs(g - grad_r).max() / max(np.abs(grad_r).max(), 1e-30)
    multipinn_b200.uninstall()
    return path, notes, el, eg, len(conds)


def main():
    names = sys.argv[1:] or sorted(d for d in os.listdir(os.path.join(REF, "examples"))
                                   if os.path.isfile(os.path.join(REF, "examples", d, "problem.py")))
    ok = bad = skipped = 0
    for name in names:
        global N_PTS
        try:
            try:
                path, notes, el, eg, k = one(name)
            except ValueError:       # composite geometries that cannot place so few points
                multipinn_b200.uninstall()
                N_PTS, keep = 400, N_PTS
                try:
                    path, notes, el, eg, k = one(name)
                finally:
                    N_PTS = keep
            flag = "OK  " if (el < 2e-5 and (np.isnan(eg) or eg < 2e-4)) else "DIFF"
            ok += flag == "OK  "
            bad += flag != "OK  "
            print(f"{flag} {name:45s} {path:5s} K={k}  loss rel err {el:.1e}  grad rel err {eg:.1e}")
            for i, n in enumerate(notes):
                print(f"       cond {i}: {n}")
        except Exception as e:   # noqa: BLE001
            skipped += 1
            msg = "".join(traceback.format_exception_only(type(e), e)).strip().splitlines()[-1]
            print(f"SKIP {name:45s} {msg[:150]}")
            if os.environ.get("SWEEP_TRACE"):
                traceback.print_exc()
            try:
                multipinn_b200.uninstall()
            except Exception:   # noqa: BLE001
                pass
        sys.stdout.flush()
    print(f"\n{ok} match, {bad} differ, {skipped} skipped (of {len(names)})")


if __name__ == "__main__":
    main()
