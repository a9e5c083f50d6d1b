"""The reference-facing Python surface (nn_module / md_module / graph_utils / train_network_* /
hack_integrator shells): same names, arguments and return conventions as the reference, results checked
against the golden vectors and the oracle.  GPU tests call through the C ABI underneath."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gamd_b200.weights import param_shapes, random_state_dict, water_bonds
from helpers import FIX, rel_err

DEV = "cuda:0"
ARGS = SimpleNamespace(use_layer_norm=True, encoding_size=128, hidden_dim=128, edge_embedding_dim=128,
                       drop_edge=False, conv_layer=4, rotate_aug=False, update_edge=False, use_part=False,
                       data_dir="", loss="mae")


def test_state_dict_keys_match_reference_layout():
    from gamd_b200.nn_module import SimpleMDNetNew, WaterMDNetNew
    m = SimpleMDNetNew(128, 3, 27.27, hidden_dim=128, conv_layer=4, edge_embedding_dim=128, drop_edge=False,
                       use_layer_norm=True)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == list(param_shapes(kind="lj").items())
    w = WaterMDNetNew(1, 128, 3, 20.0, bond=water_bonds(258), hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                      drop_edge=False, use_layer_norm=True)
    got = [(k, tuple(v.shape)) for k, v in w.state_dict().items()]
    assert got == list(param_shapes(kind="water").items())
    m.load_state_dict(random_state_dict(0, kind="lj"))          # a reference-layout checkpoint loads unchanged
    w.load_state_dict(random_state_dict(0, kind="water"))


def test_state_dict_keys_of_the_other_variants():
    """wide dynamic-box models, ``update_edge`` (per-layer ``edge_layer_norm`` registered first), ``expand_edge=False``
    (no ``edge_expand.centers``, 4 edge inputs) and BatchNorm (running statistics) keep the reference's key order."""
    from gamd_b200.nn_module import SimpleMDNetNew, WaterMDDynamicBoxNet
    for kw in (dict(encoding_size=256, hidden_dim=128, edge_embedding_dim=256, conv_layer=5),
               dict(encoding_size=128, hidden_dim=128, edge_embedding_dim=128, conv_layer=3, update_edge=True,
                    expand_edge=False),
               dict(encoding_size=768, hidden_dim=512, edge_embedding_dim=768, conv_layer=2, update_edge=True)):
        m = WaterMDDynamicBoxNet(1, kw["encoding_size"], 3, hidden_dim=kw["hidden_dim"], conv_layer=kw["conv_layer"],
                                 edge_embedding_dim=kw["edge_embedding_dim"], drop_edge=False, use_layer_norm=True,
                                 update_edge=kw.get("update_edge", False), expand_edge=kw.get("expand_edge", True))
        got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert got == list(param_shapes(kind="dynbox", use_bond=False, **kw).items())
        m.load_state_dict(random_state_dict(0, kind="dynbox", use_bond=False, **kw))
    b = SimpleMDNetNew(128, 3, 27.27, hidden_dim=128, conv_layer=4, edge_embedding_dim=128, drop_edge=False,
                       use_layer_norm=False)
    got = [(k, tuple(v.shape)) for k, v in b.state_dict().items()]
    assert got == list(param_shapes(kind="lj", use_layer_norm=False).items())
    b.load_state_dict(random_state_dict(0, kind="lj", use_layer_norm=False))


def test_model_on_cpu_fails_loudly():
    from gamd_b200 import _capi
    from gamd_b200.nn_module import SimpleMDNetNew
    m = SimpleMDNetNew(128, 3, 27.27, use_layer_norm=True, drop_edge=False).eval()
    with pytest.raises(_capi.GamdError):
        m([torch.zeros(4, 3)], [torch.zeros(2, 0, dtype=torch.long)])
    with pytest.raises(NotImplementedError):
        m.train()([torch.zeros(4, 3)], [torch.zeros(2, 0, dtype=torch.long)])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lj258_trainedstats", "lj258_batch2"])
def test_simple_mdnet_forward_golden(golden_dir, name):
    from gamd_b200.nn_module import SimpleMDNetNew
    from oracle import neighbor as onb
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m = SimpleMDNetNew(128, 3, 27.27, hidden_dim=128, conv_layer=4, edge_embedding_dim=128, drop_edge=False,
                       use_layer_norm=True)
    m.load_state_dict(random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="lj"))
    m.cuda().eval()
    pos = [torch.as_tensor(p, device=DEV) for p in g["pos"]]
    edges = [torch.as_tensor(onb.edges_bruteforce(p, 27.27, 7.5), device=DEV) for p in g["pos"]]
    for prec, tol in ((0, 2e-5), (1, 1e-4)):
        m.context(precision=prec)
        out = m(pos, edges).cpu().numpy()
        assert rel_err(out, g["force"])[0] <= tol
    # unsorted edge lists are accepted (the reference's DGL graph does not care about order)
    perm = torch.randperm(edges[0].shape[1], device=DEV)
    out2 = m([pos[0]], [edges[0][:, perm]]).cpu().numpy()
    assert rel_err(out2, g["force"][:258])[0] <= 1e-4


@pytest.mark.gpu
def test_water_mdnet_forward_golden(golden_dir):
    from gamd_b200.nn_module import WaterMDNetNew
    from oracle import neighbor as onb
    g = np.load(os.path.join(golden_dir, "tip3p774_trainedstats.npz"))
    m = WaterMDNetNew(1, 128, 3, 20.0, bond=water_bonds(258), hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                      drop_edge=False, use_layer_norm=True)
    m.load_state_dict(random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="water"))
    m.cuda().eval()
    x = torch.zeros(774, 1, device=DEV)
    x[::3] = 1.0
    pos = torch.as_tensor(g["pos"][0], device=DEV)
    edge = torch.as_tensor(onb.edges_bruteforce(g["pos"][0], 20.0, 4.2), device=DEV)
    out = m([pos], x, [edge]).cpu().numpy()
    assert rel_err(out, g["force"])[0] <= 2e-5


@pytest.mark.gpu
def test_get_neighbor_reference_order_and_masks(golden_dir):
    from gamd_b200.md_module import get_neighbor, pair_distance
    g = np.load(os.path.join(golden_dir, "get_neighbor.npz"))
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float32)
    edge, dist, norm, mb = get_neighbor(pos, 7.5, 27.27)
    assert mb is None and np.array_equal(edge.cpu().numpy(), g["lj_edge"])
    assert np.abs(norm.cpu().numpy() - g["lj_norm"]).max() <= 1e-6
    n = len(pos)
    d_all = pair_distance(pos, 27.27).cpu().numpy().reshape(n, n, 3)
    e = edge.cpu().numpy()
    assert np.abs(dist.cpu().numpy() - d_all[e[1], e[0]]).max() <= 1e-6
    # predefined_mask / bond_type are indexed in the reference's flat a*N+b order
    mask = np.zeros(n * n, bool)
    mask[::2] = True
    btype = np.arange(n * n) % 7
    e2, mb2 = get_neighbor(pos, 7.5, 27.27, return_dist=False, predefined_mask=mask, bond_type=btype)
    flat = g["lj_edge"][1].astype(np.int64) * n + g["lj_edge"][0]
    keep = mask[flat]
    assert np.array_equal(e2.cpu().numpy(), g["lj_edge"][:, keep])
    assert np.array_equal(mb2.cpu().numpy(), btype[flat[keep]])


@pytest.mark.gpu
def test_neighbor_searcher_surface():
    from gamd_b200.graph_utils import NeighborSearcher, graph_network_nbr_fn
    from oracle import neighbor as onb
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))
    s = NeighborSearcher(20.0, 4.2)
    assert not s.has_been_init
    nbr = s.init_new_neighbor_lst(pos)
    assert s.has_been_init and not nbr.did_buffer_overflow
    mask_fn = graph_network_nbr_fn(s.displacement_fn, 4.2, 774)
    mask = mask_fn(torch.as_tensor(pos, device=DEV), nbr.idx)
    # the reference's get_edge_idx (train_network_lj.py:166-185): centre ids broadcast, masked select
    center = torch.arange(774, device=DEV).view(-1, 1).expand_as(nbr.idx)[mask]
    neigh = nbr.idx[mask].long()
    ref = onb.edges_jaxmd(pos, 20.0, 4.2)
    assert np.array_equal(torch.stack([center, neigh]).cpu().numpy(), ref)
    nbr2 = s.update_neighbor_lst(pos + 0.01, nbr)
    assert np.array_equal(nbr2.edge_idx.cpu().numpy(), onb.edges_jaxmd(pos + 0.01, 20.0, 4.2))
    a, b = torch.as_tensor(pos[:5], device=DEV).float(), torch.as_tensor(pos[5:10], device=DEV).float()
    d = s.displacement_fn(a, b).cpu().numpy()
    assert np.all(np.abs(d) <= 10.0 + 1e-5)


@pytest.mark.gpu
def test_predict_forces_facades():
    from gamd_b200.train_network_lj import ParticleNetLightning as LJ
    from gamd_b200.train_network_tip3p import ParticleNetLightning as TIP3P
    from oracle import md as omd
    lj = LJ(ARGS)
    sd = random_state_dict(1, 5.2, 1.5, kind="lj")
    lj.load_state_dict({"pnet_model." + k: v for k, v in sd.items()})     # Lightning checkpoint prefix
    lj.load_training_stats(os.path.join(FIX, "scaler_lj.npz"))
    lj.cuda().eval()
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    want = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"]).predict_forces(pos)
    got = lj.predict_forces(pos)
    assert got.dtype == np.float64 and rel_err(got, want)[0] <= 1e-4

    w = TIP3P(ARGS)
    sdw = random_state_dict(4, 2.9, 0.9, kind="water")
    w.load_state_dict(sdw)
    w.load_training_stats(os.path.join(FIX, "scaler_tip3p.npz"))
    w.cuda().eval()
    sw = np.load(os.path.join(FIX, "scaler_tip3p.npz"))
    feat = torch.zeros(774, 1)
    feat[::3] = 1.0
    posw = np.load(os.path.join(FIX, "water_init_pos.npy"))
    want = omd.OracleForceField(sdw, "water", 20.0, 4.2, sw["mean"], sw["var"], bond=water_bonds(258),
                                feat=feat).predict_forces(posw)
    got = w.predict_forces(feat.cuda(), posw)
    assert rel_err(got, want)[0] <= 1e-4


@pytest.mark.gpu
def test_reference_driver_loop_nve_and_nose_hoover():
    """The loop of code/LJ/test_script/test_nosehoover.py:100-118 on the OpenMM-free hook."""
    from gamd_b200 import hack_integrator as hi
    from gamd_b200.train_network_lj import ParticleNetLightning as LJ
    from oracle import integrator as oint
    from oracle import md as omd
    lj = LJ(ARGS)
    sd = random_state_dict(1, 5.2, 1.5, kind="lj")
    lj.load_state_dict(sd)
    lj.load_training_stats(os.path.join(FIX, "scaler_lj.npz"))
    lj.cuda().eval()
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    m = np.full(258, 39.9)
    v0 = omd.maxwell_boltzmann(258, m, 100.0, 1234)
    system = hi.System(m)
    for chain in (0, 3):
        i1 = hi.HackNoseHooverIntegrator(system, 100.0, collision_frequency=25.0, chain_length=chain, timestep=0.002)
        i2 = hi.HackHalfNoseHooverIntegrator(system, 100.0, collision_frequency=25.0, chain_length=chain,
                                             timestep=0.002)
        comp = hi.CompoundIntegrator()
        comp.addIntegrator(i1)
        comp.addIntegrator(i2)
        sim = hi.Simulation(None, system, comp)
        sim.context.setPositions(pos / 10.0)
        sim.context.setVelocities(v0)
        sim.context.setPeriodicBoxSize(2.727)
        force = lj.predict_forces(sim.context.getState(getPositions=True).getPositions() * 10.0)
        # oracle twin
        x, v = pos / 10.0, v0.copy()
        f = ff.predict_forces(x * 10.0)
        st = oint.NHCState(chain, oint.KB * 100.0, 25.0, 3 * 258)
        for t in range(10):
            comp.setCurrentIntegrator(0)
            if t != 0:
                i1.copy_state_from_integrator(i2)
            i1.setPerDofVariableByName("force_last", force)
            sim.step(1)
            p = sim.context.getState(getPositions=True, enforcePeriodicBox=True).getPositions() * 10.0
            force = lj.predict_forces(p)
            comp.setCurrentIntegrator(1)
            i2.copy_state_from_integrator(i1)
            i2.setPerDofVariableByName("gnn_force", force)
            sim.step(1)
            v = oint.nhc_propagate(st, v, m, 0.002)
            x, v = oint.vv_first_half(x, v, f, m, 0.002)
            f = ff.predict_forces(x * 10.0)
            v = oint.vv_second_half(v, f, m, 0.002)
            v = oint.nhc_propagate(st, v, m, 0.002)
        got = sim.context.getState(getPositions=True, getVelocities=True, getEnergy=True)
        assert np.abs(got.getPositions() - x).max() <= 1e-7
        assert np.abs(got.getVelocities() - v).max() <= 1e-5
        assert abs(got.getKineticEnergy() - oint.kinetic_energy(v, m)) / oint.kinetic_energy(v, m) <= 1e-5
        if chain:
            assert abs(i2.getGlobalVariableByName("vxi0") - st.vxi[0]) <= 1e-6 * max(1.0, abs(st.vxi[0]))
