"""The drop-in proof: JM 19.0 lencod re-linked (GNU ld --wrap, no source edits) with jm_b200/shim/jm_wrap.c +
libjmb200.so must emit the SAME bitstream, reconstruction and syntax trace as the stock encoder on the same cfg + YUV.

Binaries (built by __graft_entry__.build() where /root/reference is mounted; they travel to the GPU box):
  oracle/_ref/lencod_ref            stock JM, every object unmodified
  jm_b200/shim/_build/lencod_jmb    the same objects + the shim; ME / sub-pel planes / transform / quant run on the GPU
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lencod_ref")
JMB = os.path.join(ROOT, "jm_b200", "shim", "_build", "lencod_jmb")
CFG = os.path.join(ROOT, "tests", "jm_cfg", "min.cfg")

needs_bins = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(JMB)),
                                reason="lencod_ref / lencod_jmb not built (needs /root/reference at build time)")


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _make_yuv(path, w, h, n, seed, fmt420=True, fade=0.0):
    from jm_b200 import synth
    frames = synth.luma_frames(w, h, n, seed=seed, motion=(3, -2))
    if fade:      # a fade to black: gives explicit weighted prediction something to find
        frames = [np.clip(np.rint(f * (1.0 - fade * i)) + 2 * i, 0, 255).astype(np.uint16) for i, f in enumerate(frames)]
    if fmt420:
        synth.write_yuv420(path, frames, textured_chroma=True)
    else:
        synth.write_yuv422(path, frames)


def _encode(exe, workdir, tag, w, h, frames, extra, env=None, bframes=0):
    args = [exe, "-d", CFG, "-p", "InputFile=input.yuv", "-p", f"SourceWidth={w}", "-p", f"SourceHeight={h}",
            "-p", f"OutputWidth={w}", "-p", f"OutputHeight={h}", "-p", f"FramesToBeEncoded={frames}",
            "-p", f"OutputFile={tag}.264", "-p", f"ReconFile={tag}_rec.yuv", "-p", f"TraceFile={tag}_trace.txt",
            "-p", "LevelIDC=40", "-p", "IntraPeriod=0", "-p", f"NumberBFrames={bframes}"]
    for kv in extra:
        args += ["-p", kv]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(args, cwd=workdir, env=e, capture_output=True, text=True, timeout=1200)
    return r


def _same_outputs(workdir, a, b):
    for suffix in (".264", "_rec.yuv", "_trace.txt"):
        assert _md5(os.path.join(workdir, a + suffix)) == _md5(os.path.join(workdir, b + suffix)), f"{suffix} differs ({a} vs {b})"


BASE = ["ProfileIDC=66", "SymbolMode=0", "RDOptimization=1", "Transform8x8Mode=0", "QPISlice=28", "QPPSlice=28"]
CONFIGS = {
    # SearchMode -1 = full_search_motion_estimation; 0 = fast_full_search_motion_estimation (JM's default)
    "full_search_baseline": BASE + ["SearchMode=-1", "SearchRange=16", "NumberReferenceFrames=2", "AdaptiveRounding=0"],
    "fast_full_search_around": BASE + ["SearchMode=0", "SearchRange=16", "NumberReferenceFrames=3", "AdaptiveRounding=1"],
    "high_8x8_cabac": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=26", "QPPSlice=27",
                       "SearchMode=-1", "SearchRange=8", "NumberReferenceFrames=1", "AdaptiveRounding=1"],
    "high_8x8_cavlc_satd8x8": ["ProfileIDC=100", "SymbolMode=0", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=30", "QPPSlice=30",
                               "SearchMode=0", "SearchRange=8", "NumberReferenceFrames=2", "AdaptiveRounding=0"],
    # BASELINE config 3 in small: EPZS (SearchMode 3) + 8x8 transform, High profile -- EPZS's predictor/threshold state machine
    # stays JM's host code, every distortion it evaluates (computeSAD / computeSATD) comes from the device
    "epzs_high_8x8": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28",
                      "SearchMode=3", "SearchRange=16", "NumberReferenceFrames=2", "AdaptiveRounding=1"],
    # the same with bin/encoder.cfg's EPZSSubPelGrid=1 (+ its pattern settings): currMB->IntPelME = EPZS_integer_motion_estimation and
    # SubPelME = EPZS_sub_pel_motion_estimation, whose WHOLE state machines run on the device -- one jmb_epzs_search call per search
    "epzs_subpelgrid_high_8x8": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28",
                                 "SearchMode=3", "SearchRange=32", "NumberReferenceFrames=3", "AdaptiveRounding=1", "EPZSSubPelGrid=1",
                                 "EPZSPattern=2", "EPZSDualRefinement=3", "EPZSFixedPredictors=3", "EPZSTemporal=1", "EPZSSpatialMem=1",
                                 "EPZSBlockType=1", "MEDistortionHPel=2", "MEDistortionQPel=2"],
    # B pictures (two between anchors, Main-style tools in High): list-1 searches and more DPB traffic go through the shim,
    # the bi-predictive refinement and direct modes stay JM's C code
    "b_frames_high": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28", "QPBSlice=30",
                      "SearchMode=0", "SearchRange=8", "NumberReferenceFrames=2", "AdaptiveRounding=1"],
    # BASELINE config 4 in small: 4:2:2 input (High 4:2:2), SATD sub-pel refinement path
    "yuv422_satd_subpel": ["ProfileIDC=122", "YUVFormat=2", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28",
                           "QPPSlice=28", "SearchMode=-1", "SearchRange=8", "NumberReferenceFrames=1", "AdaptiveRounding=1",
                           "MEDistortionHPel=2", "MEDistortionQPel=2"],
    # bi-predictive motion estimation (currMB->BiPredME / SubPelBiPredME): full_search_bipred_motion_estimation and
    # sub_pel_bipred_motion_estimation run as ONE batched device call per stage (computeBiPredSAD1 / computeBiPredSATD1)
    "b_frames_bipred_me": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28", "QPBSlice=30",
                           "SearchMode=-1", "SearchRange=8", "NumberReferenceFrames=2", "AdaptiveRounding=1", "BiPredMotionEstimation=1",
                           "BiPredMERefinements=1", "BiPredMESearchRange=8", "BiPredMESubPel=2", "BiPredSearch16x8=1", "BiPredSearch8x16=1",
                           "BiPredSearch8x8=1", "MEDistortionHPel=2", "MEDistortionQPel=2"],
    # the same with SSE as the error metric at every level (computeSSE, computeBiPredSSE1)
    "b_frames_bipred_me_sse": ["ProfileIDC=100", "SymbolMode=0", "RDOptimization=1", "Transform8x8Mode=0", "QPISlice=28", "QPPSlice=28", "QPBSlice=30",
                               "SearchMode=-1", "SearchRange=4", "NumberReferenceFrames=2", "AdaptiveRounding=0", "BiPredMotionEstimation=1",
                               "BiPredMERefinements=0", "BiPredMESearchRange=4", "BiPredMESubPel=1", "MEDistortionFPel=1", "MEDistortionHPel=1",
                               "MEDistortionQPel=1"],
    # explicit weighted prediction on a fade, weighted-reference ME: JM's own search loops run on the host and every
    # distortion (computeSADWP / computeSATDWP / computeBiPredSAD2 / computeBiPredSATD2, incl. its 8x8 branch) is the device's
    "b_frames_weighted_fade": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28", "QPBSlice=30",
                               "SearchMode=-1", "SearchRange=4", "NumberReferenceFrames=2", "AdaptiveRounding=1", "BiPredMotionEstimation=1",
                               "BiPredMERefinements=1", "BiPredMESearchRange=4", "BiPredMESubPel=2", "BiPredSearch8x8=1", "MEDistortionHPel=2",
                               "MEDistortionQPel=2", "WeightedPrediction=1", "WeightedBiprediction=1", "UseWeightedReferenceME=1"],
}


@needs_bins
@pytest.mark.parametrize("name", ["fast_full_search_around", "b_frames_bipred_me", "b_frames_weighted_fade", "yuv422_satd_subpel"])
def test_relink_is_neutral(tmp_path, name):
    """BASELINE config 0 (plumbing): with every wrapper passing through to JM's own code the re-linked encoder is
    bit-identical to the stock one -- the --wrap link of all 44 symbols itself changes nothing.  Runs on CPU."""
    w, h = 96, 80
    bframes = 2 if name.startswith("b_frames") else 0
    frames = 4 if bframes else 3
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=3, fmt420="YUVFormat=2" not in CONFIGS[name], fade=0.06 if "fade" in name else 0.0)
    r1 = _encode(REF, tmp_path, "ref", w, h, frames, CONFIGS[name], bframes=bframes)
    r2 = _encode(JMB, tmp_path, "pt", w, h, frames, CONFIGS[name], env={"JMB_SHIM": "passthrough"}, bframes=bframes)
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr[-500:], r2.stderr[-500:])
    _same_outputs(tmp_path, "ref", "pt")


@needs_bins
def test_no_gpu_is_a_loud_error(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    _make_yuv(tmp_path / "input.yuv", 64, 48, 2, seed=4)
    r = _encode(JMB, tmp_path, "g", 64, 48, 2, CONFIGS["full_search_baseline"])
    assert r.returncode != 0 and "no CPU path" in r.stderr


@pytest.mark.gpu
@needs_bins
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_bitstream_identical_to_stock_jm(tmp_path, name):
    """Motion vectors, quantised coefficients and the emitted bitstream, bit-exact against the JM CPU encoder."""
    w, h, frames = 96, 80, 4
    bframes = 2 if name.startswith("b_frames") else 0
    if bframes:
        frames = 4 if name in ("b_frames_bipred_me_sse", "b_frames_weighted_fade") else 7      # (these two run JM's own search loops, one device call per candidate)
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=11, fmt420="YUVFormat=2" not in CONFIGS[name], fade=0.06 if "fade" in name else 0.0)
    r1 = _encode(REF, tmp_path, "ref", w, h, frames, CONFIGS[name], bframes=bframes)
    r2 = _encode(JMB, tmp_path, "gpu", w, h, frames, CONFIGS[name], env={"JMB_SHIM_VERBOSE": "1"}, bframes=bframes)
    assert r1.returncode == 0, r1.stderr[-800:]
    assert r2.returncode == 0, r2.stderr[-800:]
    line = [l for l in r2.stderr.splitlines() if l.startswith("[jmb shim]")]
    assert line, "the shim did not report: was the GPU path used?"
    counts = dict(zip(line[0].split()[2::2][:9], [int(x) for x in line[0].split()[3::2][:9]]))
    assert counts["planes"] >= max(1, (frames - 1) // (bframes + 1)) and counts["quant4"] + counts["quant8"] > 0, line[0]
    if "EPZSSubPelGrid=1" in CONFIGS[name]:
        # whole searches as single device calls (EPZS_integer_motion_estimation / EPZS_sub_pel_motion_estimation); JM's
        # sub-macroblock and bi-predictive EPZS variants still ask for one distortion at a time
        assert counts["full"] > 0 and counts["subpel"] > 0, line[0]
    elif "SearchMode=3" in CONFIGS[name] or "UseWeightedReferenceME=1" in CONFIGS[name]:
        assert counts["dist"] > 0, line[0]                       # EPZS / weighted-reference ME: distortion oracle
    else:
        assert counts["full"] + counts["fastfull"] > 0 and counts["subpel"] > 0, line[0]
    if "BiPredMotionEstimation=1" in CONFIGS[name]:
        assert counts["dist"] > 0, line[0]                       # the bi-predictive candidates went through jmb_dist_ex
        assert int(line[0].split("bipred/weighted distortions")[1].split()[0]) > 0, line[0]
    _same_outputs(tmp_path, "ref", "gpu")


VERIFY_CONFIGS = {
    "baseline_4x4_cavlc": CONFIGS["full_search_baseline"],
    "high_8x8_cavlc": CONFIGS["high_8x8_cavlc_satd8x8"],
    "high_8x8_cabac_2refs": ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=24", "QPPSlice=25",
                             "SearchMode=0", "SearchRange=8", "NumberReferenceFrames=2", "AdaptiveRounding=0"],
}


@pytest.mark.gpu
@needs_bins
@pytest.mark.parametrize("name", sorted(VERIFY_CONFIGS))
def test_device_luma_residual_coding_matches_jm_in_the_live_encoder(tmp_path, name):
    """Differential pin of the whole-macroblock device path (jmb_luma_residual_coding: prediction .. thresholding ..
    reconstruction) against the REAL luma_residual_coding: with JMB_SHIM_VERIFY=1 the shim recodes every inter macroblock
    JM residual-codes (every RD candidate) on the device from JM's own state and compares levels, cbp, cbp_blk and the
    reconstruction; the first difference aborts the encoder."""
    w, h, frames = 96, 80, 3
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=17)
    r = _encode(JMB, tmp_path, "v", w, h, frames, VERIFY_CONFIGS[name], env={"JMB_SHIM_VERIFY": "1"})
    assert r.returncode == 0, r.stderr[-800:]
    line = [l for l in r.stderr.splitlines() if "luma_residual_coding verified on" in l]
    assert line, r.stderr[-400:]
    assert int(line[0].split("verified on")[1].split()[0]) > 100, line[0]
    assert int(line[0].split("chroma_residual_coding verified on")[1].split()[0].rstrip(";")) > 100, line[0]      # 4:2:0 chroma: prediction, 2x2 DC path, AC, thresholds
    # ... and every coded picture was deblocked on the device (k_deblock) AND by JM's own DeblockFrame, sample for sample the same
    assert int(line[0].split("DeblockFrame verified on")[1].split()[0]) == frames, line[0]


@pytest.mark.gpu
@needs_bins
@pytest.mark.parametrize("name", sorted(VERIFY_CONFIGS))
def test_luma_residual_coding_answered_by_the_device_gives_the_same_bitstream(tmp_path, name):
    """JMB_SHIM_RC=device: for inter macroblocks of P slices JM's luma_residual_coding is NOT run -- one device call per RD
    candidate (jmb_luma_residual_coding: prediction .. transform .. quantisation .. reconstruction .. thresholding) and its
    levels, cbp, cbp_blk and reconstruction written into JM's state (cofAC, currMB, enc_picture).  Mode decision, entropy coding
    and everything after see exactly what JM's own function would have left: bitstream, reconstruction and trace are identical."""
    w, h, frames = 96, 80, 3
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=19)
    r1 = _encode(REF, tmp_path, "ref", w, h, frames, VERIFY_CONFIGS[name])
    r2 = _encode(JMB, tmp_path, "gpu", w, h, frames, VERIFY_CONFIGS[name], env={"JMB_SHIM_RC": "device", "JMB_SHIM_VERBOSE": "1"})
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr[-500:], r2.stderr[-500:])
    _same_outputs(tmp_path, "ref", "gpu")
    line = [l for l in r2.stderr.splitlines() if l.startswith("[jmb shim]")][0]
    assert int(line.split("luma_residual_coding answered by the device")[1].split()[0]) > 100, line


@pytest.mark.gpu
@needs_bins
def test_device_chroma_residual_coding_422_matches_jm_in_the_live_encoder(tmp_path):
    """The same differential pin for 4:2:2 (BASELINE config 4's chroma): hadamard4x2 + quant_dc4x2 at qp + 3, eight AC blocks per
    component, 1/4-sample vertical chroma motion."""
    w, h, frames = 96, 80, 3
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=18, fmt420=False)
    cfg = ["ProfileIDC=122", "YUVFormat=2", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=0", "QPISlice=26", "QPPSlice=27",
           "SearchMode=0", "SearchRange=8", "NumberReferenceFrames=2", "AdaptiveRounding=0", "MEDistortionHPel=2", "MEDistortionQPel=2"]
    r = _encode(JMB, tmp_path, "v", w, h, frames, cfg, env={"JMB_SHIM_VERIFY": "1"})
    assert r.returncode == 0, r.stderr[-800:]
    line = [l for l in r.stderr.splitlines() if "chroma_residual_coding verified on" in l]
    assert line and int(line[0].split("chroma_residual_coding verified on")[1].split()[0].rstrip(";")) > 100, r.stderr[-400:]
    assert int(line[0].split("DeblockFrame verified on")[1].split()[0]) == frames, line[0]      # 4:2:2 deblocking: sixteen chroma rows, four horizontal chroma edges


FIXTURES = os.path.join(ROOT, "oracle", "_ref", "fixtures")


@pytest.mark.gpu
@needs_bins
@pytest.mark.skipif(not os.path.isdir(FIXTURES), reason="reference fixtures not copied (make -C oracle ref)")
@pytest.mark.parametrize("cfg", ["encoder_baseline.cfg", "encoder.cfg", "encoder_yuv422.cfg"])
def test_bundled_configurations(tmp_path, cfg):
    """BASELINE configs[0] and its siblings exactly as shipped: JM's own cfg files on JM's own foreman clips (copied into
    the git-ignored oracle/_ref/fixtures by the oracle Makefile).  encoder_baseline.cfg = Baseline, fast full search +-32,
    5 references, adaptive rounding; encoder.cfg = High, EPZS + HME, 7 B pictures, bi-predictive ME, 8x8 transform, CABAC;
    encoder_yuv422.cfg = High 4:2:2.  Stock encoder vs re-linked encoder: same bitstream, reconstruction and trace."""
    import shutil
    for f in os.listdir(FIXTURES):
        shutil.copy(os.path.join(FIXTURES, f), tmp_path / f)
    import time
    outs, secs = {}, {}
    for tag, exe, env in (("ref", REF, {}), ("gpu", JMB, {"JMB_SHIM_VERBOSE": "1"})):
        e = dict(os.environ); e.update(env)
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-d", cfg, "-p", f"OutputFile={tag}.264", "-p", f"ReconFile={tag}_rec.yuv", "-p", f"TraceFile={tag}_trace.txt"],
                           cwd=tmp_path, env=e, capture_output=True, text=True, timeout=1500)
        secs[tag] = time.perf_counter() - t0
        assert r.returncode == 0, (tag, r.stderr[-800:])
        outs[tag] = r
    assert any(l.startswith("[jmb shim]") for l in outs["gpu"].stderr.splitlines())
    _same_outputs(tmp_path, "ref", "gpu")
    # wall time of the two encoders (process start and CUDA context creation included), kept with the run's artefacts
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        me = {t: [l.strip() for l in outs[t].stdout.splitlines() if "Total ME time" in l or "Total encoding time" in l] for t in outs}
        with open(os.path.join(log, "dropin_times.txt"), "a") as f:
            f.write(f"{cfg}: lencod_ref {secs['ref']:.2f} s, lencod_jmb {secs['gpu']:.2f} s | ref {me['ref']} | jmb {me['gpu']}\n")


@pytest.mark.gpu
@needs_bins
@pytest.mark.parametrize("name", ["fast_full_search_around", "full_search_baseline"])
def test_searches_run_on_resident_surfaces_with_device_argmin(tmp_path, name):
    """No arg-min on the host: the macroblock's SAD surfaces stay on the device (jmb_mb_surfaces) and every partition's search is
    a device arg-min over them (jmb_mb_search); the sub-pel refinement BlockMotionSearch asks for next comes out of the same
    device call.  Same bitstream as stock JM."""
    w, h, frames = 96, 80, 3
    _make_yuv(tmp_path / "input.yuv", w, h, frames, seed=12)
    r1 = _encode(REF, tmp_path, "ref", w, h, frames, CONFIGS[name])
    r2 = _encode(JMB, tmp_path, "gpu", w, h, frames, CONFIGS[name], env={"JMB_SHIM_VERBOSE": "1"})
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr[-500:], r2.stderr[-500:])
    _same_outputs(tmp_path, "ref", "gpu")
    line = [l for l in r2.stderr.splitlines() if l.startswith("[jmb shim]")][0]
    served = int(line.split("served by the integer search's call")[1].split()[0]); builds = int(line.split("surface builds")[1].split()[0])
    searches = int(line.split("full")[1].split()[0]) + int(line.split("fastfull")[1].split()[0])
    assert builds > 0 and served > 0.9 * searches, line
    # ... and most searches were answered AHEAD of JM's call (jmb_mb_chain: the searches of a macroblock region in one device call,
    # each answer handed out only if JM arrives with the predictor and centre the device derived)
    calls = int(line.split("chain calls")[1].split()[0]); ahead = int(line.split("searches answered ahead")[1].split()[0])
    assert calls > 0 and ahead + calls > 0.5 * searches, line
    # the same with the run-ahead switched off: one device call per search, the same bitstream
    r3 = _encode(JMB, tmp_path, "gpu1", w, h, frames, CONFIGS[name], env={"JMB_SHIM_VERBOSE": "1", "JMB_SHIM_CHAIN": "0"})
    assert r3.returncode == 0, r3.stderr[-500:]
    _same_outputs(tmp_path, "ref", "gpu1")
    assert "chain calls 0 " in [l for l in r3.stderr.splitlines() if l.startswith("[jmb shim]")][0]
