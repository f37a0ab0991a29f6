//! Drop-in for the reference crate `feature_extraction` (feature_extraction/src/lib.rs): the same public items with
//! the same signatures; the bodies marshal OpenCV containers across the C ABI of `libdunk_b200.so`
//! (include/dunk_b200.h) instead of calling OpenCV.  Citations are lines of the reference file.
use dunk_b200_sys as sys;
use opencv::{
    core::{DMatch, KeyPoint, Mat, Point2f, Scalar, Vector, CV_8UC1},
    features2d::DrawMatchesFlags,
    imgcodecs,
    prelude::*,
    Error,
};

pub const MAX_POINTS_SHIFT: i32 = 18; // lib.rs:12
pub const MAX_POINTS: i32 = (1 << MAX_POINTS_SHIFT) - 1; // lib.rs:13
const DESC_BYTES: usize = sys::DUNK_DESC_BYTES as usize;

/// lib.rs:15-18
pub struct ExtractedKeyPoint {
    keypoints: Vector<KeyPoint>,
    descriptors: Mat,
}

/// lib.rs:20-31
#[derive(Debug)]
pub struct DbKeypoints {
    pub x_coord: f32,
    pub y_coord: f32,
    pub size: f32,
    pub angle: f32,
    pub response: f32,
    pub octave: i32,
    pub class_id: i32,
    pub descriptor: Vec<u8>,
    pub image_id: i32,
}

impl ExtractedKeyPoint {
    /// lib.rs:33-59 — one DB row per keypoint, the descriptor row copied out of the N x 61 matrix
    pub fn to_db_type(&self, image_id: i32) -> Vec<DbKeypoints> {
        let bytes = self.descriptors.data_bytes().unwrap_or(&[]);
        self.keypoints
            .iter()
            .enumerate()
            .map(|(i, kp)| DbKeypoints {
                x_coord: kp.pt().x,
                y_coord: kp.pt().y,
                size: kp.size(),
                angle: kp.angle(),
                response: kp.response(),
                octave: kp.octave(),
                class_id: kp.class_id(),
                descriptor: bytes[i * DESC_BYTES..(i + 1) * DESC_BYTES].to_vec(),
                image_id,
            })
            .collect()
    }

    /// additive accessors (the reference keeps the fields private and reads them inside the crate)
    pub fn keypoints(&self) -> &Vector<KeyPoint> {
        &self.keypoints
    }
    pub fn descriptors(&self) -> &Mat {
        &self.descriptors
    }
}

fn check(rc: i32) -> Result<(), Error> {
    if rc == 0 {
        Ok(())
    } else {
        // the library's status codes ARE OpenCV's (-215 StsAssert, -211 StsOutOfRange, -28 StsVecLengthErr, ...)
        Err(Error::new(rc, sys::last_error()))
    }
}

/// (data pointer, rows, cols, channels, row stride in bytes) of an 8-bit Mat
fn mat_u8_view(m: &Mat) -> Result<(*const u8, i32, i32, i32, i32), Error> {
    if m.empty() {
        return Ok((std::ptr::null(), 0, 0, 1, 0));
    }
    if m.depth() != opencv::core::CV_8U {
        return Err(Error::new(sys::DUNK_ERR_ASSERT, "8-bit image / descriptor matrix expected".to_string()));
    }
    let step = m.step1(0)? as i32; // elements per row == bytes per row for CV_8U
    Ok((m.data(), m.rows(), m.cols(), m.channels(), step))
}

/// contiguous N x desc_bytes view of a descriptor Mat (copies only if the Mat is not continuous)
fn desc_rows(m: &Mat) -> Result<(std::borrow::Cow<'_, [u8]>, i32, i32), Error> {
    let (_, rows, cols, ch, _) = mat_u8_view(m)?;
    let width = cols * ch;
    if rows == 0 {
        return Ok((std::borrow::Cow::Borrowed(&[]), 0, width));
    }
    if m.is_continuous() {
        Ok((std::borrow::Cow::Borrowed(m.data_bytes()?), rows, width))
    } else {
        Ok((std::borrow::Cow::Owned(m.try_clone()?.data_bytes()?.to_vec()), rows, width))
    }
}

/// lib.rs:61-92 — AKAZE(DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, DIFF_PM_G2, max_points or 2^18 - 1).detectAndCompute
pub fn akaze_keypoint_descriptor_extraction_def(img: &Mat, max_points: Option<i32>) -> Result<ExtractedKeyPoint, Error> {
    let (data, rows, cols, channels, stride) = mat_u8_view(img)?;
    // output capacity: the library's raw-candidate default (w*h/32, at least 2048), far above real keypoint counts
    let cap = ((rows as i64 * cols as i64 / 32).clamp(2048, 1 << 20)) as usize;
    let mut kps = vec![sys::DunkKeyPoint { x: 0.0, y: 0.0, size: 0.0, angle: 0.0, response: 0.0, octave: 0, class_id: 0 }; cap];
    let mut desc = vec![0u8; cap * DESC_BYTES];
    let mut n = 0i32;
    // SAFETY: `data` addresses rows x stride readable bytes owned by `img` for the duration of the call; the output
    // buffers hold `cap` elements / rows as promised by the `cap` argument
    check(unsafe {
        sys::dunk_akaze_extract(sys::ctx(), data, rows, cols, channels, stride, max_points.unwrap_or(MAX_POINTS), kps.as_mut_ptr(),
                                desc.as_mut_ptr(), cap as i32, &mut n)
    })?;
    let n = n as usize;
    let mut keypoints = Vector::<KeyPoint>::with_capacity(n);
    for k in &kps[..n] {
        keypoints.push(KeyPoint::new_coords(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id)?);
    }
    let descriptors = if n == 0 {
        Mat::default()
    } else {
        let mut m = Mat::new_rows_cols_with_default(n as i32, DESC_BYTES as i32, CV_8UC1, Scalar::all(0.0))?;
        m.data_bytes_mut()?.copy_from_slice(&desc[..n * DESC_BYTES]);
        m
    };
    Ok(ExtractedKeyPoint { keypoints, descriptors })
}

fn dmatches(raw: &[sys::DunkDMatch]) -> Vector<DMatch> {
    let mut v = Vector::<DMatch>::with_capacity(raw.len());
    for m in raw {
        v.push(DMatch { query_idx: m.query_idx, train_idx: m.train_idx, img_idx: m.img_idx, distance: m.distance });
    }
    v
}

/// lib.rs:94-114 — BFMatcher(NORM_HAMMING).knnMatch(k) + Lowe ratio `d0 < d1 * filter_strength` (f32, strict).
/// Fewer than 2 train rows -> Err(-211), what the reference's `i.get(1)?` yields (lib.rs:108).
pub fn get_knn_matches(origin_desc: &Mat, target_desc: &Mat, k: i32, filter_strength: f32) -> Result<Vector<DMatch>, Error> {
    let (q, nq, wq) = desc_rows(origin_desc)?;
    let (t, nt, wt) = desc_rows(target_desc)?;
    if nq > 0 && nt > 0 && wq != wt {
        return Err(Error::new(sys::DUNK_ERR_ASSERT, "query and train descriptors differ in width".to_string()));
    }
    let mut out = vec![sys::DunkDMatch { query_idx: 0, train_idx: 0, img_idx: 0, distance: 0.0 }; nq.max(1) as usize];
    let mut n = 0i32;
    // SAFETY: q / t are nq x wq and nt x wt contiguous byte rows; `out` holds nq records
    check(unsafe {
        sys::dunk_knn_match_hamming(sys::ctx(), q.as_ptr(), nq, t.as_ptr(), nt as i64, wq.max(wt), k, filter_strength, out.as_mut_ptr(),
                                    out.len() as i32, &mut n)
    })?;
    Ok(dmatches(&out[..n as usize]))
}

/// lib.rs:116-126 — BFMatcher(NORM_HAMMING, crossCheck = true).match
pub fn get_bruteforce_matches(origin_desc: &Mat, target_desc: &Mat) -> Result<Vector<DMatch>, Error> {
    let (q, nq, wq) = desc_rows(origin_desc)?;
    let (t, nt, wt) = desc_rows(target_desc)?;
    let mut out = vec![sys::DunkDMatch { query_idx: 0, train_idx: 0, img_idx: 0, distance: 0.0 }; nq.max(1) as usize];
    let mut n = 0i32;
    // SAFETY: as in get_knn_matches
    check(unsafe {
        sys::dunk_match_crosscheck_hamming(sys::ctx(), q.as_ptr(), nq, t.as_ptr(), nt as i64, wq.max(wt), out.as_mut_ptr(), out.len() as i32,
                                           &mut n)
    })?;
    Ok(dmatches(&out[..n as usize]))
}

/// lib.rs:128-155 — drawing + file output: not on the hot path, stays on OpenCV
pub fn export_matches(img1: &Mat, img1_keypoints: &Vector<KeyPoint>, img2: &Mat, img2_keypoints: &Vector<KeyPoint>,
                      matches: &Vector<DMatch>, export_location: &str) -> Result<(), Error> {
    let mut canvas = Mat::default();
    opencv::features2d::draw_matches(img1, img1_keypoints, img2, img2_keypoints, matches, &mut canvas, Scalar::all(-1.0), Scalar::all(-1.0),
                                     &Vector::<i8>::new(), DrawMatchesFlags::NOT_DRAW_SINGLE_POINTS)?;
    imgcodecs::imwrite(export_location, &canvas, &Vector::new())?;
    Ok(())
}

/// lib.rs:157-159 — file input: stays on OpenCV
pub fn get_mat_from_dir(img_dir: &str) -> Result<Mat, Error> {
    imgcodecs::imread(img_dir, imgcodecs::IMREAD_COLOR)
}

/// lib.rs:161-180.  The reference indexes image-1 keypoints by `m.img_idx` (always 0, :169) and converts the image-1
/// list into BOTH outputs (:176-177); this keeps the signature and implements the evident intent — `query_idx`
/// addresses image 1, `train_idx` image 2 — as documented in DESIGN.md ("Deviations").
pub fn get_points_from_matches(img1_keypoints: &Vector<KeyPoint>, img2_keypoints: &Vector<KeyPoint>,
                               matches: &Vector<DMatch>) -> Result<(Vector<Point2f>, Vector<Point2f>), Error> {
    let mut p1 = Vector::<Point2f>::with_capacity(matches.len());
    let mut p2 = Vector::<Point2f>::with_capacity(matches.len());
    for m in matches {
        p1.push(img1_keypoints.get(usize::try_from(m.query_idx).map_err(|_| Error::new(-211, "negative query_idx".to_string()))?)?.pt());
        p2.push(img2_keypoints.get(usize::try_from(m.train_idx).map_err(|_| Error::new(-211, "negative train_idx".to_string()))?)?.pt());
    }
    Ok((p1, p2))
}
