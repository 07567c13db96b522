"""-m gpu: THE DROP-IN PROOF.  oracle/_ref/libref_dropin.so is the reference's unmodified src/Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc,
KeyFrameDatabase.cc (+ DBoW2 containers) compiled with this repository's include/ORBextractor.h, ORBmatcher.h and ORBVocabulary.h
swapped in for the reference's three headers (-DORB_B200_WITH_REFERENCE_TYPES; recipe: oracle/Makefile) and linked against
liborb_b200.so.  The reference's call sites are untouched:
    (*mpORBextractor)(im, cv::Mat(), mvKeys, mDescriptors)              src/Frame.cc:60
    ORBmatcher matcher(0.9, true); matcher.SearchByProjection(mCurrentFrame, mLastFrame, 15)     src/Tracking.cc:596-605
    ORBmatcher::DescriptorDistance(vDescriptors[i], vDescriptors[j])    src/MapPoint.cc:224
    mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)      src/Frame.cc:285
    mpVoc->score(F->mBowVec, pKFi->GetBowVector())                      src/KeyFrameDatabase.cc:247
Every scenario of tests/test_ref_build.py (which pins the ORACLE against the reference's own ORBextractor.cc / ORBmatcher.cc) is re-run
with the drop-in library in the reference's place, and the TrackWithMotionModel call sequence is run on both libraries side by side."""
import numpy as np
import pytest

import test_ref_build as T
from oracle import pydropin, pyref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pydropin.available(), reason="oracle/_ref/libref_dropin.so not built (needs /root/reference at build time)")]


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.fixture(autouse=True)
def _dropin_in_place_of_reference(monkeypatch):
    monkeypatch.setattr(T, "pyref", pydropin)
    assert pydropin.flavour().startswith("dropin")


def _cases(fn):
    """the parametrize lists of a test function of test_ref_build.py as keyword dictionaries"""
    out = [{}]
    for mark in getattr(fn, "pytestmark", []):
        if mark.name != "parametrize":
            continue
        names = [n.strip() for n in mark.args[0].split(",")]
        out = [dict(o, **dict(zip(names, v if len(names) > 1 else (v,)))) for o in out for v in mark.args[1]]
    return out


def _run_all(fn, **fixtures):
    cases = _cases(fn)
    for kw in cases:
        fn(**fixtures, **kw)
    return len(cases)


def test_extractor_call_site(po):
    """ORBextractor::operator()(cv::InputArray, cv::InputArray, vector<cv::KeyPoint>&, cv::OutputArray): 8 configurations x 6 frames"""
    assert _run_all(T.test_extractor_equals_reference, po=po) == 8
    T.test_extractor_degenerate_frames(po)


def test_frame_constructor_unmodified(po):
    """Frame::Frame (src/Frame.cc:56-128) as compiled from the reference: extraction on the GPU, then ITS UndistortKeyPoints, bounds, grid"""
    assert _run_all(T.test_frame_constructor_equals_oracle, po=po) == 3


def test_descriptor_distance_and_distinctive(po):
    T.test_descriptor_distance(po)
    assert _run_all(T.test_distinctive_descriptor, po=po) == 3          # src/MapPoint.cc:185-250 calling ORBmatcher::DescriptorDistance


def test_search_by_projection_frame_frame(po):
    assert _run_all(T.test_search_by_projection_frame_frame, po=po) == 5


def test_search_by_projection_mappoints(po):
    assert _run_all(T.test_search_by_projection_mappoints, po=po) == 3


def test_window_search_and_projection_window(po):
    assert _run_all(T.test_window_search, po=po) == 3
    assert _run_all(T.test_search_by_projection_window, po=po) == 3


def test_search_for_initialization(po, pkg):
    assert _run_all(T.test_search_for_initialization, po=po) == 3
    T.test_search_for_initialization_steal_chains(po, pkg)


def test_search_by_bow(po, pkg):
    assert _run_all(T.test_search_by_bow_keyframe_frame, po=po, pkg=pkg) == 4
    assert _run_all(T.test_search_by_bow_keyframe_keyframe, po=po, pkg=pkg) == 3


def test_search_for_triangulation(po, pkg):
    assert _run_all(T.test_search_for_triangulation, po=po, pkg=pkg) == 4


def test_back_end_searches_with_host_preamble(po):
    """the five methods whose pose / Sim3 projection is host arithmetic in the reference: restated with the same cv:: expressions in
    include/orb_b200_reftypes.h, scoring on the GPU, graph bookkeeping (Replace / AddObservation) on the host in the reference's order"""
    assert _run_all(T.test_search_by_projection_keyframe, po=po) == 3          # SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist)
    assert _run_all(T.test_search_by_projection_sim3, po=po) == 3              # SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)
    assert _run_all(T.test_fuse, po=po) == 3                                   # Fuse(KeyFrame*, vpMapPoints, th)
    assert _run_all(T.test_search_by_sim3, po=po) == 3                         # SearchBySim3
    assert _run_all(T.test_fuse_sim3, po=po) == 2                              # Fuse(KeyFrame*, Scw, vpPoints, th)


def test_track_with_motion_model_equals_reference(po):
    """src/Tracking.cc:594-606 on both libraries: the reference's ORBmatcher.cc and the drop-in, same frames, same velocity"""
    for shape, nf in (((240, 320), 500), ((480, 752), 1000), ((376, 1241), 2000)):
        (ka, da), (kb, db), cam, has, outl, xyz, Tcw = T._scene(po, shape[0], shape[1], nf, 9100 + nf)
        res = []
        for mod in (pyref, pydropin):
            last = mod.RefFrame(ka, da, *cam).set_mappoints(has, xyz, outl)
            cur = mod.RefFrame(kb, db, *cam)
            res.append(mod.track_with_motion_model(cur, last, Tcw))           # mLastFrame.mTcw = I, so mVelocity is the new pose
        (n_ref, m_ref), (n_gpu, m_gpu) = res
        assert n_ref > 20 and n_ref == n_gpu and np.array_equal(m_ref, m_gpu)
        ocur, olast = po.OracleFrame(kb, db, *cam), po.OracleFrame(ka, da, *cam)
        rn, rm = po.search_by_projection(ocur, olast, has, outl, xyz, Tcw, 15.0, True)
        assert rn == n_gpu and np.array_equal(np.asarray(rm), m_gpu)


@pytest.mark.parametrize("k,L,levelsup,prune,order", [(10, 3, 1, 0.0, "bfs"), (4, 5, 4, 0.15, "dfs"), (3, 6, 4, 0.1, "bfs")])
def test_vocabulary_call_sites(po, tmp_path, k, L, levelsup, prune, order):
    """ORBVocabulary(): loadFromTextFile, transform(vector<cv::Mat>, BowVector&, FeatureVector&, levelsup), transform(feature), score"""
    from orbslam_jpminipc_b200 import synth
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=k * 10 + L, prune_frac=prune, order=order)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, k, L, parent, desc, weight, trailing_newline=False)
    gv, ov = pydropin.RefVocabulary(path), po.OracleVocabulary(path=path)
    assert gv.nwords == ov.nwords
    rng = np.random.default_rng(L)
    leaves = np.nonzero(~np.isin(np.arange(len(parent)), parent))[0]
    bows = []
    for n in (1, 37, 1000):
        feats = desc[rng.choice(leaves, n)] ^ np.packbits((rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1)
        (gw, gvv), (gn_, gs, gi) = gv.transform(feats, levelsup)
        (ow, ovv), (on_, os_, oi) = ov.transform(feats, levelsup)
        assert np.array_equal(gw, ow) and np.array_equal(gvv.view(np.uint64), ovv.view(np.uint64))     # BowVector doubles bit for bit
        assert np.array_equal(gn_, on_) and np.array_equal(gs, os_) and np.array_equal(gi, oi)
        w, _, _ = ov.transform_features(feats[:20], 0)
        assert [gv.word(f) for f in feats[:20]] == list(w)
        bows.append((ow, ovv))
    for a in bows:
        for b in bows:      # every caller stores the score in a float (src/KeyFrameDatabase.cc:140,247; src/LoopClosing.cc:139): the shim returns that float
            assert np.float32(gv.score(a, b)) == np.float32(po.bow_score_l1(a, b))


def test_relocalisation_candidates_through_keyframe_database(po, tmp_path):
    """KeyFrameDatabase::DetectRelocalisationCandidates as compiled from the reference, scoring through the GPU vocabulary"""
    T.test_relocalisation_candidate_scoring(po, tmp_path)
