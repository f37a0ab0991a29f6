"""CPU: the Rust shim (rust/, uncompiled here: the image has no cargo / rustc) stays complete — the generated sys crate
declares every symbol of include/dunk_b200.h and is up to date, and the three shim crates define every public item of
the reference crates that SURVEY 8a keeps (names taken from the reference's signatures)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(*p):
    return open(os.path.join(ROOT, *p)).read()


def test_sys_crate_is_generated_from_the_header_and_complete():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    src = re.sub(r"/\*.*?\*/", "", read("include", "dunk_b200.h"), flags=re.S)
    syms = sorted(set(re.findall(r"\b(dunk_[a-z0-9_]+)\s*\(", src)))
    rs = read("rust", "dunk-b200-sys", "src", "lib.rs")
    assert len(syms) >= 90 and not [s for s in syms if f"pub fn {s}(" not in rs]
    for struct in ("DunkKeyPoint", "DunkDMatch", "DunkTop2", "DunkImage", "DunkRowFilter", "DunkRegistration", "DunkPoseConfig", "DunkPose",
                   "DunkPipelineView"):
        assert f"pub struct {struct} " in rs


def test_shim_crates_define_the_kept_public_items():
    fe = read("rust", "feature_extraction", "src", "lib.rs")
    for item in ("pub const MAX_POINTS_SHIFT", "pub const MAX_POINTS", "pub struct ExtractedKeyPoint", "pub struct DbKeypoints",
                 "pub fn to_db_type(&self, image_id: i32) -> Vec<DbKeypoints>",
                 "pub fn akaze_keypoint_descriptor_extraction_def(img: &Mat, max_points: Option<i32>) -> Result<ExtractedKeyPoint, Error>",
                 "pub fn get_knn_matches(origin_desc: &Mat, target_desc: &Mat, k: i32, filter_strength: f32) -> Result<Vector<DMatch>, Error>",
                 "pub fn get_bruteforce_matches(origin_desc: &Mat, target_desc: &Mat) -> Result<Vector<DMatch>, Error>",
                 "pub fn export_matches(", "pub fn get_mat_from_dir(img_dir: &str) -> Result<Mat, Error>", "pub fn get_points_from_matches("):
        assert item in fe, item
    hg = read("rust", "homographier", "src", "homographier", "mod.rs")
    for item in ("pub trait PixelElemType", "pub struct BGRA", "pub enum HomographyMethod", "RHO = 16", "pub enum MatError", "Jagged",
                 "pub struct PNPRANSACSolution", "pub struct ImgObjCorrespondence", "pub struct Cmat<T>", "pub fn from_2d_slice(",
                 "pub fn new(mat: Mat) -> Result<Self, MatError>", "pub fn imread_checked(", "pub fn at_2d(&self, row: i32, col: i32)",
                 "pub fn zeros(rows: i32, cols: i32)", "impl<T> ToInputArray for Cmat<T>", "impl<T> ToOutputArray for Cmat<T>",
                 "pub fn raster_to_mat(pixels: &[RGBA8], w: i32, h: i32) -> Result<Cmat<Vec4b>, MatError>",
                 "pub fn find_homography_mat(", "pub fn warp_image_perspective<T: DataType>(", "pub fn pnp_solver_ransac("):
        assert item in hg, item
    assert "pub mod homographier;" in read("rust", "homographier", "src", "lib.rs")
    kp = read("rust", "feature_database", "src", "keypointdb.rs")
    for fn in ("create_keypoint", "read_keypoint_from_id", "read_keypoints_from_image_id", "read_keypoints_from_lod",
               "read_keypoints_from_coordinates", "delete_keypoint"):
        assert kp.count(f"fn {fn}(") == 2, fn                     # trait declaration + implementation
    im = read("rust", "feature_database", "src", "imagedb.rs")
    for fn in ("create_image", "read_image_from_id", "find_images_from_dimensions", "find_images_from_lod", "delete_image"):
        assert im.count(f"fn {fn}(") == 2, fn
    el = read("rust", "feature_database", "src", "elevationdb.rs")
    for fn in ("create_geotransform", "get_world_coordinates", "add_elevation_data", "get_elevation"):
        assert f"pub fn {fn}(" in el, fn
    md = read("rust", "feature_database", "src", "models.rs")
    for st in ("pub struct Image", "pub struct InsertImage<'a>", "pub struct Keypoint", "pub struct InsertKeypoint<'a>"):
        assert st in md, st


def test_every_ffi_call_of_the_shims_exists_in_the_sys_crate():
    rs = read("rust", "dunk-b200-sys", "src", "lib.rs")
    declared = set(re.findall(r"pub fn (dunk_\w+)\(", rs)) | {"last_error", "ctx"}
    for crate, files in (("feature_extraction", ["lib.rs"]), ("homographier", ["homographier/mod.rs"]),
                         ("feature_database", ["lib.rs", "imagedb.rs", "keypointdb.rs", "elevationdb.rs"])):
        for f in files:
            used = set(re.findall(r"sys::(dunk_\w+|last_error|ctx)\(", read("rust", crate, "src", *f.split("/"))))
            assert used and used <= declared, (crate, f, used - declared)
