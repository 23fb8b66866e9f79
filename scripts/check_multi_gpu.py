"""Multi-GPU parity check (run under torchrun on N GPUs): the column-sharded psi — with the all-gather fused into
the kernel as peer stores, and with the separate NCCL all-gather — must equal the single-GPU host-API matrix
bit for bit on EVERY rank, and the first error must be the same global pair on every rank."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

import pharmsol_b200 as ps
from benches import harness as H, workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for name, nsub, nspp, kw in [("c1", 40, 1003, {}), ("c2", 64, 2500, dict(solver=ps.OdeSolver.Dopri5, tol=1e-6)), ("c3", 16, 777, {}),
                             ("c5", 6, 130, dict(particles=96))]:
    w = W.make(name, nsub=nsub, nspp=nspp, **({"particles": kw["particles"]} if "particles" in kw else {}))
    eq, data, ems = H.product_objects(w, device=local)
    if "solver" in kw:
        eq.with_solver(kw["solver"]).with_tolerances(kw["tol"], kw["tol"])
    if "particles" in kw:      # SDE streams are keyed by the global pair index: sharding must not change psi
        eq.with_particles(kw["particles"]).with_mode(ps.SdeMode.ParticleFilter).with_seed(99)
    ref = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)          # whole matrix on this GPU, host API
    # fused peer stores | one NCCL all-gather | phased: 7/8 of the columns gathered while the last 1/8 computes
    for peer, overlap, gmode in ((True, False, "auto"), (False, False, "auto"), (False, True, "auto"), (True, False, "push")):
        job = ps.ResidentPsi(eq, data, w["support_points"], ems, device=dev, peer_stores=("auto" if gmode == "push" else peer), gather_overlap=overlap, gather=gmode)
        fused = getattr(job.sharded, "peer_ptrs", None) is not None
        ok = ok and (gmode != "push" or job.gather_mode == "push")
        for _ in range(2):      # twice: the second step runs over the first one's status / buffers
            job.step()
        psi = job.finish().cpu().numpy()
        same = np.array_equal(psi, ref, equal_nan=True)
        ok = ok and same and (fused == peer) and (len(job.ranges) == (2 if overlap else 1))
        print(f"rank {rank} {name} gather={job.gather_mode} peer_stores={peer} overlap={overlap} fused={fused} equal_to_single_gpu={same} err={getattr(job.sharded, 'peer_error', None)}", flush=True)
        dist.barrier()
# first error propagates to every rank with its global pair index
eq = ps.Equation.from_dsl("name = twocpt\nkind = analytical\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = cp\nbolus(iv) -> central\n"
                          "structure = two_compartments\nout(cp) = central / v ~ continuous()\n")
ops = [("bolus", 0.0, 100.0, "iv"), ("observation", 1.0, 50.0, "cp")]
data = ps.Data([ps.Subject(f"s{i}", ops) for i in range(5)])
ems = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
for peer, overlap, bad in ((True, False, 333), (False, False, 333), (False, True, 390), (False, True, 20)):
    spp = np.tile(np.array([[0.1, 3.0, 1.0, 1.0]]), (400, 1))
    spp[bad] = [1.0, -3.0, 1.5, 1.0]      # with two phases column 390 falls in the tail phase, column 20 in the head phase
    job = ps.ResidentPsi(eq, data, spp, ems, device=dev, peer_stores=peer, gather_overlap=overlap)
    job.step()
    try:
        job.finish()
        ok = False
        print(f"rank {rank}: expected an error", flush=True)
    except ps.PharmsolError as e:
        good = e.code == 12 and e.pair == 0 + bad * 5
        ok = ok and good
        print(f"rank {rank} error propagation peer_stores={peer} overlap={overlap}: code {e.code} pair {e.pair} ok={good}", flush=True)
    dist.barrier()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
