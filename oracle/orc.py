"""ctypes face of the CPU oracle (oracle/libhaf_oracle.so) and of the in-place compiled reference
(oracle/_ref/libhaf_ref.so, svm-scale, svm-predict).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libhaf_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libhaf_ref.so")
REFSERVER_SO = os.path.join(REF_DIR, "libhaf_refserver.so")


def build(ref: bool = True) -> None:
    """(Re)build the checker.  `make ref` is a no-op where /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def ref_available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libhaf_ref.so", "svm-scale", "svm-predict"))


class OrcRequest(C.Structure):
    _fields_ = [("center", C.c_double * 3), ("area_len_x", C.c_float), ("area_len_y", C.c_float),
                ("approach", C.c_double * 3), ("gripper_opening_width", C.c_int), ("return_only_best", C.c_int),
                ("graspval_top", C.c_int), ("roll_limit", C.c_int)]


class OrcBest(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("roll", C.c_int), ("tilt", C.c_int), ("topval", C.c_int),
                ("eval", C.c_int), ("roll_rad", C.c_float), ("rolls_done", C.c_int), ("n_windows", C.c_long)]

    def astuple(self):
        return (self.row, self.col, self.roll, self.tilt, self.topval)


def make_request(center=(0.0, 0.0, 0.0), area=(32.0, 44.0), approach=(0.0, 0.0, 1.0), width=1,
                 return_only_best=0, graspval_top=119, roll_limit=0) -> OrcRequest:
    rq = OrcRequest()
    rq.center[:] = center
    rq.area_len_x, rq.area_len_y = area
    rq.approach[:] = approach
    rq.gripper_opening_width = width
    rq.return_only_best = return_only_best
    rq.graspval_top = graspval_top
    rq.roll_limit = roll_limit
    return rq


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The restatement.  Holds the parsed feature table, range file and (optionally) a model."""

    def __init__(self, features_path: str, range_path: str, model_path: str | None = None,
                 nr_features_without_shaf: int = 302):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.L = C.CDLL(ORACLE_SO)
        for name in ("orc_features_load", "orc_range_load", "orc_svm_load"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_char_p]
        for name in ("orc_features_free", "orc_range_free", "orc_svm_free"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_features_count.argtypes = [C.c_void_p]
        L.orc_range_max_index.argtypes = [C.c_void_p]
        L.orc_svm_total_sv.argtypes = [C.c_void_p]
        L.orc_svm_gamma.argtypes = [C.c_void_p]
        L.orc_svm_gamma.restype = C.c_double
        L.orc_svm_rho.argtypes = [C.c_void_p]
        L.orc_svm_rho.restype = C.c_double
        L.orc_text4.restype = C.c_double
        L.orc_text4.argtypes = [C.c_float]
        L.orc_text6.restype = C.c_double
        L.orc_text6.argtypes = [C.c_double]
        L.orc_generate_grid.restype = C.c_long
        self.nshaf = nr_features_without_shaf
        self.feat = L.orc_features_load(features_path.encode())
        if not self.feat:
            raise FileNotFoundError(features_path)
        self.range = L.orc_range_load(range_path.encode())
        if not self.range:
            raise FileNotFoundError(range_path)
        self.F = L.orc_features_count(C.c_void_p(self.feat))
        self.svm = None
        if model_path:
            self.load_model(model_path)

    def load_model(self, model_path: str):
        if self.svm:
            self.L.orc_svm_free(C.c_void_p(self.svm))
        self.svm = self.L.orc_svm_load(model_path.encode())
        if not self.svm:
            raise ValueError("oracle cannot parse model " + model_path)

    # ---- stage functions -------------------------------------------------------------
    def feature_table(self):
        reg = np.zeros((self.F, 16), np.int32)
        w = np.zeros((self.F, 4), np.float32)
        for i in range(self.F):
            self.L.orc_features_get(C.c_void_p(self.feat), i, _fp(reg[i]), _fp(w[i]))
        return reg, w

    def normalize_approach(self, av):
        a = (C.c_double * 3)(*av)
        o = (C.c_double * 3)()
        self.L.orc_normalize_approach(a, o)
        return tuple(o)

    def build_transform(self, center, av_norm, width, roll, roll_step_deg=15):
        M = np.zeros(16, np.float32)
        self.L.orc_build_transform((C.c_double * 3)(*center), (C.c_double * 3)(*av_norm), int(width), int(roll),
                                   int(roll_step_deg), _fp(M))
        return M

    def generate_grid(self, xyz: np.ndarray, M: np.ndarray, G: int = 56, want_cells: bool = False):
        xyz = np.ascontiguousarray(xyz, np.float32)
        heights = np.zeros((G, G), np.float32)
        cells = np.zeros(len(xyz), np.int32) if want_cells else None
        clamped = self.L.orc_generate_grid(_fp(xyz), C.c_size_t(len(xyz)), C.c_size_t(xyz.strides[0]), _fp(M), G,
                                           _fp(heights), _fp(cells) if want_cells else None)
        return (heights, cells, clamped) if want_cells else heights

    def calc_intimage(self, heights: np.ndarray):
        G = heights.shape[0]
        integral = np.zeros((G + 1, G + 1), np.float32)
        self.L.orc_calc_intimage(_fp(np.ascontiguousarray(heights, np.float32)), G, _fp(integral))
        return integral

    def pnt_in_box(self, integral: np.ndarray, roll: int, area=(32, 44), roll_step_deg=15, boxrot=0.0):
        G = integral.shape[0] - 1
        mask = np.zeros((G, G), np.uint8)
        self.L.orc_pnt_in_box(_fp(np.ascontiguousarray(integral, np.float32)), G, int(roll), int(roll_step_deg),
                              int(area[0]), int(area[1]), C.c_float(boxrot), _fp(mask))
        return mask

    def calc_featurevectors(self, integral: np.ndarray, mask: np.ndarray):
        G = mask.shape[0]
        integral = np.ascontiguousarray(integral, np.float32)
        mask = np.ascontiguousarray(mask, np.uint8)
        W = self.L.orc_calc_featurevectors(C.c_void_p(self.feat), _fp(integral), G, _fp(mask), self.nshaf, None,
                                           None, 0)
        feats = np.zeros((W, self.F), np.float32)
        rc = np.zeros((W, 2), np.int32)
        self.L.orc_calc_featurevectors(C.c_void_p(self.feat), _fp(integral), G, _fp(mask), self.nshaf, _fp(feats),
                                       _fp(rc), W)
        return feats, rc

    def featurevalues(self, patch: np.ndarray):
        out = np.zeros(self.F, np.float32)
        self.L.orc_calc_featurevalues(C.c_void_p(self.feat), _fp(np.ascontiguousarray(patch, np.float32)), self.nshaf,
                                      _fp(out))
        return out

    def scale(self, feats: np.ndarray, emulate_text: bool = True):
        W, F = feats.shape
        self.L.orc_scale_dim.argtypes = [C.c_void_p, C.c_int]
        dim = self.L.orc_scale_dim(C.c_void_p(self.range), F)
        scaled = np.zeros((W, dim), np.float64)
        self.L.orc_scale(C.c_void_p(self.range), _fp(np.ascontiguousarray(feats, np.float32)), W, F,
                         int(emulate_text), _fp(scaled))
        return scaled

    def text4(self, v):
        return self.L.orc_text4(C.c_float(v))

    def text6(self, v):
        return self.L.orc_text6(C.c_double(v))

    def svm_decision(self, scaled: np.ndarray):
        W, dim = scaled.shape
        dec = np.zeros(W, np.float64)
        lab = np.zeros(W, np.int32)
        self.L.orc_svm_decision(C.c_void_p(self.svm), _fp(np.ascontiguousarray(scaled, np.float64)), W, dim, _fp(dec),
                                _fp(lab))
        return dec, lab

    # ---- probability estimates (svm-predict -b 1; server.cpp:831-841) ------------------------
    def check_probability_model(self):
        self.L.orc_svm_check_probability_model.argtypes = [C.c_void_p]
        return bool(self.L.orc_svm_check_probability_model(C.c_void_p(self.svm)))

    def svm_predict_probability(self, scaled: np.ndarray):
        """(labels [W] float64, prob_estimates [W][2] in the model's label order)"""
        W, dim = scaled.shape
        lab = np.zeros(max(W, 1), np.float64)
        pr = np.zeros((max(W, 1), 2), np.float64)
        rc = self.L.orc_svm_predict_probability(C.c_void_p(self.svm), _fp(np.ascontiguousarray(scaled, np.float64)), W, dim, _fp(lab), _fp(pr))
        if rc != 0:
            raise ValueError("Model does not support probabiliy estimates")
        return lab[:W], pr[:W]

    def format_probability_output(self, labels: np.ndarray, probs: np.ndarray) -> bytes:
        """the file svm-predict -b 1 writes for these predictions"""
        W = len(labels)
        cap = 64 + 96 * W
        buf = C.create_string_buffer(cap)
        self.L.orc_format_probability_output.restype = C.c_long
        self.L.orc_format_probability_output.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        n = self.L.orc_format_probability_output(C.c_void_p(self.svm), _fp(np.ascontiguousarray(labels, np.float64)),
                                                 _fp(np.ascontiguousarray(probs, np.float64)), W, buf, cap)
        assert n >= 0
        return buf.raw[:n]

    def show_predicted_gps_prob(self, output_text: bytes, mask: np.ndarray):
        """show_predicted_gps(roll, tilt, svm_with_probability=true) on the text of /tmp/output_calc_gp.txt:
        (graspsgrid, graspseval, (row, col, topval))"""
        G = mask.shape[0]
        grid = np.zeros((G, G), np.float32)
        ev = np.zeros((G, G), np.float32)
        top = np.zeros(3, np.int32)
        self.L.orc_show_predicted_gps_prob.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.orc_show_predicted_gps_prob(output_text, _fp(np.ascontiguousarray(mask, np.uint8)), G, _fp(grid), _fp(ev), _fp(top))
        return grid, ev, tuple(int(t) for t in top)

    def search_prob(self, xyz: np.ndarray, rq: OrcRequest, G: int = 56, roll_step_deg: int = 15, roll_max_deg: int = 190):
        """loop_control (server.cpp:335-402) with svm_with_probability = true: per-roll stage functions, the -b 1 output text,
        show_predicted_gps' probability branch; strict > across rolls (:953), early exit (:362-365)."""
        R = roll_max_deg // roll_step_deg
        if rq.roll_limit > 0:
            R = min(R, rq.roll_limit)
        av = self.normalize_approach(tuple(rq.approach))
        area = (int(rq.area_len_x), int(rq.area_len_y))
        best = (-1, -1, -1, -1, -1000)
        out = dict(graspseval=[], graspsgrid=[], per_roll_top=[], mask=[], probs=[], text=[])
        for roll in range(R):
            if rq.return_only_best and best[4] >= rq.graspval_top:
                break
            M = self.build_transform(tuple(rq.center), av, rq.gripper_opening_width, roll, roll_step_deg)
            integral = self.calc_intimage(self.generate_grid(xyz, M, G))
            mask = self.pnt_in_box(integral, roll, area, roll_step_deg)
            feats, _ = self.calc_featurevectors(integral, mask)
            scaled = self.scale(feats) if len(feats) else np.zeros((0, 1))
            lab, pr = self.svm_predict_probability(scaled)
            text = self.format_probability_output(lab, pr)
            grid, ev, top = self.show_predicted_gps_prob(text, mask)
            if top[2] > best[4]:
                best = (top[0], top[1], roll, 0, top[2])
            out["graspseval"].append(ev); out["graspsgrid"].append(grid); out["per_roll_top"].append(top)
            out["mask"].append(mask); out["probs"].append(pr); out["text"].append(text)
        out["best"] = best
        out["graspseval"] = np.array(out["graspseval"]); out["per_roll_top"] = np.array(out["per_roll_top"], np.int32)
        return out

    def show_predicted_gps(self, labels: np.ndarray, mask: np.ndarray):
        G = mask.shape[0]
        ev = np.zeros((G, G), np.float32)
        top = np.zeros(3, np.int32)
        first = np.zeros(2, np.int32)
        self.L.orc_show_predicted_gps(_fp(np.ascontiguousarray(labels, np.int32)),
                                      _fp(np.ascontiguousarray(mask, np.uint8)), G, _fp(ev), _fp(top), _fp(first))
        return ev, tuple(int(t) for t in top), tuple(int(t) for t in first)

    def transform_gp_in_wcs(self, rq: OrcRequest, heights_best: np.ndarray, row: int, col: int, roll: int, roll_step_deg: int = 15):
        """a16: (gp1 xyz, gp2 xyz, averaged xyz, approach xyz, roll rad) as 13 doubles."""
        out = np.zeros(13, np.float64)
        G = heights_best.shape[0]
        self.L.orc_transform_gp_in_wcs((C.c_double * 3)(*rq.center), (C.c_double * 3)(*rq.approach), int(rq.gripper_opening_width),
                                       int(roll_step_deg), G, _fp(np.ascontiguousarray(heights_best, np.float32)), int(row), int(col),
                                       int(roll), _fp(out))
        return out

    def search(self, xyz: np.ndarray, rq: OrcRequest, G: int = 56, roll_step_deg: int = 15, roll_max_deg: int = 190,
               emulate_text: bool = True, full: bool = True):
        xyz = np.ascontiguousarray(xyz, np.float32)
        R = roll_max_deg // roll_step_deg
        best = OrcBest()
        out = {}
        if full:
            out["heights"] = np.zeros((R, G, G), np.float32)
            out["integral"] = np.zeros((R, G + 1, G + 1), np.float32)
            out["mask"] = np.zeros((R, G, G), np.uint8)
            out["graspseval"] = np.zeros((R, G, G), np.float32)
            out["per_roll_top"] = np.full((R, 3), -1, np.int32)
            cap = R * (G - 14) * (G - 14)
            out["dec"] = np.zeros(cap, np.float64)
        else:
            cap = 0
        rc = self.L.orc_search(_fp(xyz), C.c_size_t(len(xyz)), C.c_size_t(xyz.strides[0]), C.byref(rq),
                               C.c_void_p(self.feat), C.c_void_p(self.range), C.c_void_p(self.svm), G, roll_step_deg,
                               roll_max_deg, self.nshaf, int(emulate_text), C.byref(best),
                               _fp(out["heights"]) if full else None, _fp(out["integral"]) if full else None,
                               _fp(out["mask"]) if full else None, _fp(out["graspseval"]) if full else None,
                               _fp(out["per_roll_top"]) if full else None, _fp(out["dec"]) if full else None,
                               C.c_long(cap))
        assert rc == 0
        if full:
            out["dec"] = out["dec"][:best.n_windows]
        out["best"] = best
        return out


class Ref:
    """The reference's own compiled code (oracle/_ref)."""

    def __init__(self, features_path: str, range_path: str, model_path: str | None = None,
                 nr_features_without_shaf: int = 302):
        if not ref_available():
            raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
        L = self.L = C.CDLL(REF_SO)
        L.ref_features_new.restype = C.c_void_p
        L.ref_features_new.argtypes = [C.c_char_p]
        L.ref_svm_load.restype = C.c_void_p
        L.ref_svm_load.argtypes = [C.c_char_p]
        L.ref_svm_predict.restype = C.c_double
        self.nshaf = nr_features_without_shaf
        self.range_path = range_path
        self.model_path = model_path
        self.feat = L.ref_features_new(features_path.encode())
        self.F = L.ref_features_count(C.c_void_p(self.feat))
        self.svm = L.ref_svm_load(model_path.encode()) if model_path else None

    def feature_table(self):
        reg = np.zeros((self.F, 16), np.int32)
        w = np.zeros((self.F, 4), np.float64)
        for i in range(self.F):
            self.L.ref_features_get(C.c_void_p(self.feat), i, _fp(reg[i]), _fp(w[i]))
        return reg, w

    def featurevalues(self, patch: np.ndarray):
        out = np.zeros(self.F, np.float32)
        self.L.ref_features_calc(C.c_void_p(self.feat), _fp(np.ascontiguousarray(patch, np.float32)), self.nshaf,
                                 _fp(out))
        return out

    def svm_predict(self, x_dense: np.ndarray):
        x = np.ascontiguousarray(x_dense, np.float64)
        dec = C.c_double()
        lab = self.L.ref_svm_predict(C.c_void_p(self.svm), _fp(x), len(x), C.byref(dec))
        return dec.value, int(lab)

    def roll_file_exact(self, integral: np.ndarray, mask: np.ndarray, workdir: str | None = None):
        """One roll exactly as server.cpp:616-656 + :754-800 do it: the reference class appends one text
        line per window, then the reference svm-scale and svm-predict executables run as child processes
        on the files.  Returns (labels [W] as parsed by server.cpp:843, scaled text lines)."""
        G = mask.shape[0]
        own = workdir is None
        wd = tempfile.mkdtemp(prefix="hafref_") if own else workdir
        f_feat = os.path.join(wd, "features.txt")
        f_scale = os.path.join(wd, "features.txt.scale")
        f_out = os.path.join(wd, "output_calc_gp.txt")
        open(f_feat, "w").close()  # server.cpp:632 truncates
        integral = np.ascontiguousarray(integral, np.float32)
        fb = f_feat.encode()
        for row in range(G - 14):
            for col in range(G - 14):
                if not mask[row + 7, col + 7]:
                    continue
                patch = np.ascontiguousarray(integral[row:row + 15, col:col + 15])
                self.L.ref_features_write(C.c_void_p(self.feat), _fp(patch), fb, self.nshaf)
        with open(f_scale, "w") as so:
            subprocess.run([os.path.join(REF_DIR, "svm-scale"), "-r", self.range_path, f_feat], stdout=so,
                           stderr=subprocess.DEVNULL, check=True)
        subprocess.run([os.path.join(REF_DIR, "svm-predict"), f_scale, self.model_path, f_out],
                       stdout=subprocess.DEVNULL, check=True)
        with open(f_out) as fh:
            labels = [int(ln[:2]) for ln in fh.read().split("\n") if ln]  # atoi(line.substr(0,2)), :843
        with open(f_scale) as fh:
            scaled_lines = fh.read().split("\n")[:-1]
        if own:
            for f in (f_feat, f_scale, f_out):
                os.remove(f)
            os.rmdir(wd)
        return np.array(labels, np.int32), scaled_lines


def refserver_available() -> bool:
    return os.path.exists(REFSERVER_SO) and os.path.exists(os.path.join(REF_DIR, "pkg", "libsvm-3.12", "svm-predict"))


class RefServer:
    """The reference's own action server class CCalc_Grasppoints (src/calc_grasppoints_action_server.cpp compiled in place,
    unmodified, against stand-in ROS / PCL / Eigen / OpenCV headers: oracle/server_shim.cpp, oracle/stub_server).  One goal
    runs read_pc_cb -> loop_control -> ... -> transform_gp_in_wcs_and_publish as the reference wrote them, child processes
    and /tmp/features.txt included (so: one RefServer goal at a time per machine)."""

    def __init__(self, features_path: str, range_path: str, model_path: str):
        if not refserver_available():
            raise RuntimeError("oracle/_ref/libhaf_refserver.so not built (run `make -C oracle ref` where /root/reference exists)")
        L = self.L = C.CDLL(REFSERVER_SO)
        L.refsrv_new.restype = C.c_void_p
        L.refsrv_new.argtypes = [C.c_char_p] * 4
        L.refsrv_free.argtypes = [C.c_void_p]
        L.refsrv_run_goal.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                      C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 9
        L.refsrv_transform_of_roll.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.G, self.R = L.refsrv_grid(), L.refsrv_rolls()
        self.h = L.refsrv_new(features_path.encode(), range_path.encode(), model_path.encode(), os.path.join(REF_DIR, "pkg").encode())

    def close(self):
        if self.h:
            self.L.refsrv_free(C.c_void_p(self.h))
            self.h = None

    def run_goal(self, xyz, center=(0.0, 0.0, 0.0), area=(32.0, 44.0), approach=(0.0, 0.0, 1.0), width=1, only_best=0,
                 max_time_s=1e6, per_roll=True):
        xyz = np.ascontiguousarray(xyz, np.float32)
        G, R = self.G, self.R
        out = dict(best=np.zeros(5, np.int32), grasp=np.zeros(14, np.float64), heights=np.zeros((R, G, G), np.float32),
                   integral=np.zeros((R, G + 1, G + 1), np.float32), mask=np.zeros((R, G, G), np.uint8),
                   per_roll_top=np.full((R, 3), -1, np.int32), eval_pos=np.zeros((R, G, G), np.float32),
                   eval_seen=np.zeros((R, G, G), np.uint8), M_last=np.zeros(16, np.float32))
        c = np.array(center, np.float64)
        a = np.array(approach, np.float64)
        rc = self.L.refsrv_run_goal(C.c_void_p(self.h), _fp(xyz), len(xyz), xyz.strides[0] if len(xyz) else 12, _fp(c), area[0], area[1], _fp(a),
                                    int(width), int(only_best), float(max_time_s), int(per_roll), _fp(out["best"]), _fp(out["grasp"]),
                                    _fp(out["heights"]), _fp(out["integral"]), _fp(out["mask"]), _fp(out["per_roll_top"]),
                                    _fp(out["eval_pos"]), _fp(out["eval_seen"]), _fp(out["M_last"]))
        assert rc == 0
        return out

    def prob_rolls(self, xyz, prob_model_path):
        """show_predicted_gps(roll, 0, true) for every roll of the goal last run on this server (same cloud), with the reference's
        own svm-predict -b 1 as the child process: (per_roll_top [R][3], eval_pos [R][G][G], eval_seen [R][G][G])"""
        xyz = np.ascontiguousarray(xyz, np.float32)
        G, R = self.G, self.R
        top = np.full((R, 3), -1, np.int32)
        pos = np.zeros((R, G, G), np.float32)
        seen = np.zeros((R, G, G), np.uint8)
        cmd = "%s -b 1 /tmp/features.txt.scale %s /tmp/output_calc_gp.txt > /dev/null" % (os.path.join(REF_DIR, "svm-predict"), prob_model_path)
        self.L.refsrv_prob_rolls.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.L.refsrv_prob_rolls(C.c_void_p(self.h), _fp(xyz), len(xyz), xyz.strides[0] if len(xyz) else 12, cmd.encode(), _fp(top), _fp(pos), _fp(seen))
        assert rc == 0
        return top, pos, seen

    def transform_of_roll(self, roll):
        M = np.zeros(16, np.float32)
        self.L.refsrv_transform_of_roll(C.c_void_p(self.h), int(roll), _fp(M))
        return M
