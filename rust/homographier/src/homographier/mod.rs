//! Drop-in for homographier/src/homographier/mod.rs: the public items keep their names, signatures and error
//! behaviour; `find_homography_mat`, `pnp_solver_ransac`, `warp_image_perspective` and `raster_to_mat` run on the
//! GPU through `libdunk_b200.so`.  `Cmat<T>` / `MatError` are the reference's checked-matrix contract (API surface,
//! no arithmetic).  Citations are lines of the reference file.
use std::marker::PhantomData;

use dunk_b200_sys as sys;
use opencv::{
    calib3d::SolvePnPMethod,
    core::{Point2d, Point2f, Point3d, Scalar, Size2i, ToInputArray, ToOutputArray, Vec4b, CV_8UC4},
    prelude::*,
    Error,
};
use rgb::RGBA8;

/// mod.rs:14-23
pub trait PixelElemType {
    fn to_cv_const(&self) -> i32;
}
pub struct BGRA;
impl PixelElemType for BGRA {
    fn to_cv_const(&self) -> i32 {
        CV_8UC4
    }
}

/// mod.rs:25-31 — the values are OpenCV's (and `DunkHomographyMethod`'s)
#[derive(Clone, Copy)]
pub enum HomographyMethod {
    Default = 0,
    LMEDS = 4,
    RANSAC = 8,
    RHO = 16,
}

/// mod.rs:33-44
#[non_exhaustive]
#[derive(Debug)]
pub enum MatError {
    /// an OpenCV-style error: the library's status codes are OpenCV's (-215, -211, -28, ...)
    Opencv(opencv::Error),
    /// the matrix holds no data
    Empty,
    /// rows of differing length
    Jagged,
    Unknown,
}

/// mod.rs:46-51
#[derive(Debug)]
pub struct PNPRANSACSolution {
    pub rvec: Cmat<f64>,
    pub tvec: Cmat<f64>,
    pub inliers: Cmat<i32>,
}

/// mod.rs:53-65 — a 3-D object point and the image point it projects to
pub struct ImgObjCorrespondence {
    pub obj_point: Point3d,
    pub img_point: Point2d,
}

impl ImgObjCorrespondence {
    pub fn new(obj_point: Point3d, img_point: Point2d) -> Self {
        Self { obj_point, img_point }
    }
}

/// mod.rs:71-75 — a Mat known to hold data of element type T
#[derive(Debug)]
pub struct Cmat<T> {
    pub mat: Mat,
    _marker: PhantomData<T>,
}

impl<T> Cmat<T> {
    fn non_empty(mat: Mat) -> Result<Self, MatError> {
        // Mat::dims() is 0 only for an empty matrix (mod.rs:85-91)
        if mat.dims() == 0 {
            Err(MatError::Empty)
        } else {
            Ok(Cmat { mat, _marker: PhantomData })
        }
    }

    /// mod.rs:93-100
    pub fn from_2d_slice(slice: &[impl AsRef<[T]>]) -> Result<Self, MatError>
    where
        T: DataType,
    {
        Cmat::new(Mat::from_slice_2d::<T>(slice).map_err(MatError::Opencv)?)
    }

    fn check(&self) -> Result<(), MatError> {
        if self.mat.dims() == 0 {
            Err(MatError::Empty)
        } else {
            Ok(())
        }
    }
}

impl<T: DataType> Cmat<T> {
    /// mod.rs:114-119 — a type mismatch is reported as `Empty`, as the reference does
    pub fn new(mat: Mat) -> Result<Self, MatError> {
        if T::opencv_type() == mat.typ() {
            Cmat::non_empty(mat)
        } else {
            Err(MatError::Empty)
        }
    }

    /// mod.rs:121-124 — file input, stays on OpenCV
    pub fn imread_checked(filename: &str, flags: i32) -> Result<Self, MatError> {
        Cmat::new(opencv::imgcodecs::imread(filename, flags).map_err(MatError::Opencv)?)
    }

    /// mod.rs:129-137 — the reference compares `row` with the WIDTH and `col` with the HEIGHT (strict `>`) before
    /// Mat::at_2d's own check; kept as is (reference test cmat_at_2d_works, :605-625)
    pub fn at_2d(&self, row: i32, col: i32) -> Result<&T, MatError> {
        let size = self.mat.size().map_err(|_| MatError::Unknown)?;
        if row > size.width || col > size.height {
            return Err(MatError::Opencv(Error::new(sys::DUNK_ERR_OUT_OF_RANGE, "")));
        }
        self.mat.at_2d::<T>(row, col).map_err(MatError::Opencv)
    }

    /// mod.rs:139-145
    pub fn zeros(rows: i32, cols: i32) -> Result<Cmat<T>, MatError> {
        let m = Mat::zeros(rows, cols, T::opencv_type()).map_err(MatError::Opencv)?.to_mat().map_err(MatError::Opencv)?;
        Cmat::new(m)
    }
}

fn as_cv_error(e: MatError) -> opencv::Error {
    match e {
        MatError::Opencv(inner) => inner,
        _ => opencv::Error { code: -2, message: "unknown error".into() },
    }
}

/// mod.rs:148-171
impl<T> ToInputArray for Cmat<T> {
    fn input_array(&self) -> opencv::Result<opencv::core::_InputArray> {
        self.check().map_err(as_cv_error)?;
        self.mat.input_array()
    }
}
impl<T> ToOutputArray for Cmat<T> {
    fn output_array(&mut self) -> opencv::Result<opencv::core::_OutputArray> {
        self.check().map_err(as_cv_error)?;
        self.mat.output_array()
    }
}

fn status(rc: i32) -> Result<(), MatError> {
    if rc == 0 {
        Ok(())
    } else {
        Err(MatError::Opencv(Error::new(rc, sys::last_error())))
    }
}

/// mod.rs:183-197 (+ raster_1d_to_2d :199-216, rbga8_to_vec4b :218-220): `w * h` RGBA8 pixels -> BGRA `Cmat<Vec4b>`.
/// The per-pixel swizzle runs on the device (`dunk_raster_to_mat`); a length mismatch is `MatError::Unknown`.
pub fn raster_to_mat(pixels: &[RGBA8], w: i32, h: i32) -> Result<Cmat<Vec4b>, MatError> {
    if w < 0 || h < 0 || pixels.len() != (w as usize) * (h as usize) {
        return Err(MatError::Unknown);
    }
    if pixels.is_empty() {
        return Err(MatError::Empty);
    }
    let mut mat = Mat::new_rows_cols_with_default(h, w, CV_8UC4, Scalar::all(0.0)).map_err(MatError::Opencv)?;
    // SAFETY: RGBA8 is 4 packed bytes; source and destination both hold w * h pixels
    status(unsafe { sys::dunk_raster_to_mat(sys::ctx(), pixels.as_ptr() as *const u8, w, h, mat.data_mut()) })?;
    Cmat::new(mat)
}

/// mod.rs:231-259 — findHomography(input, reference, mask, method or Default, threshold or 3.0) with the 5-argument
/// overload's maxIters 2000 / confidence 0.995.  The mask is returned for RANSAC and LMEDS only (:253-257).
/// Fewer than 4 pairs -> Opencv(-28); no model -> `MatError::Empty` (OpenCV returns an empty Mat, `Cmat::new` rejects it).
pub fn find_homography_mat(input: &[Point2f], reference: &[Point2f], method: Option<HomographyMethod>,
                           reproj_threshold: Option<f64>) -> Result<(Cmat<f64>, Option<Cmat<u8>>), MatError> {
    if input.len() != reference.len() {
        return Err(MatError::Opencv(Error::new(sys::DUNK_ERR_ASSERT, "input and reference differ in length".to_string())));
    }
    let n = input.len() as i32;
    let mut h = [0f64; 9];
    let mut mask = vec![0u8; input.len().max(1)];
    let mut found = 0i32;
    // SAFETY: Point2f is two packed f32; both slices hold n points; mask holds n bytes
    status(unsafe {
        sys::dunk_find_homography(sys::ctx(), input.as_ptr() as *const f32, reference.as_ptr() as *const f32, n,
                                  method.unwrap_or(HomographyMethod::Default) as i32, reproj_threshold.unwrap_or(3.0), h.as_mut_ptr(),
                                  mask.as_mut_ptr(), &mut found)
    })?;
    if found == 0 {
        return Err(MatError::Empty);
    }
    let hm = Cmat::<f64>::from_2d_slice(&[&h[0..3], &h[3..6], &h[6..9]])?;
    let out_mask = match method {
        Some(HomographyMethod::RANSAC) | Some(HomographyMethod::LMEDS) => {
            let rows: Vec<[u8; 1]> = mask[..input.len()].iter().map(|&b| [b]).collect();
            Some(Cmat::<u8>::from_2d_slice(&rows)?)
        }
        _ => None,
    };
    Ok((hm, out_mask))
}

/// mod.rs:271-300 — warpPerspective(src, m, size or src.size(), INTER_LINEAR, BORDER_CONSTANT, Scalar(1,1,1,1)) for
/// 8-bit images of 1..4 channels, bit-exact with OpenCV 4.13 (1/32-pixel grid, 15-bit weights).
pub fn warp_image_perspective<T: DataType>(src: &Cmat<T>, m: &Cmat<f64>, size: Option<Size2i>) -> Result<Cmat<T>, MatError> {
    let src_size = src.mat.size().map_err(|_| MatError::Unknown)?;
    let size = size.unwrap_or(src_size);
    if m.mat.rows() != 3 || m.mat.cols() != 3 {
        return Err(MatError::Opencv(Error::new(sys::DUNK_ERR_ASSERT, "m must be 3x3".to_string())));
    }
    if src.mat.depth() != opencv::core::CV_8U {
        return Err(MatError::Opencv(Error::new(sys::DUNK_ERR_BAD_ARG, "8-bit images only".to_string())));
    }
    let mut mm = [0f64; 9];
    for r in 0..3 {
        for c in 0..3 {
            mm[(r * 3 + c) as usize] = *m.mat.at_2d::<f64>(r, c).map_err(MatError::Opencv)?;
        }
    }
    let channels = src.mat.channels();
    let mut dst = Mat::new_rows_cols_with_default(size.height, size.width, src.mat.typ(), Scalar::new(1.0, 1.0, 1.0, 1.0))
        .map_err(MatError::Opencv)?;
    let stride = src.mat.step1(0).map_err(MatError::Opencv)? as i32;
    // SAFETY: src holds rows x stride bytes; dst was allocated for out_rows x out_cols x channels bytes (continuous)
    status(unsafe {
        sys::dunk_warp_perspective(sys::ctx(), src.mat.data(), src_size.height, src_size.width, channels, stride, mm.as_ptr(), size.height,
                                   size.width, std::ptr::null(), dst.data_mut())
    })?;
    Cmat::<T>::new(dst)
}

/// mod.rs:320-369 — solvePnPRansac(obj, img, K, zeros(4,1), rvec, tvec, false, iter_count, reproj_thres, confidence,
/// inliers, method or SOLVEPNP_EPNP).  `dist_coeffs` is accepted and ignored exactly as in the reference (shadowed by
/// zeros at :344).  `Ok(None)`: no pose found (:367).  Fewer than 4 correspondences -> Opencv(-215) (test :627-638).
pub fn pnp_solver_ransac(point_correspondences: &[ImgObjCorrespondence], camera_intrinsic: &Cmat<f64>, iter_count: i32,
                         reproj_thres: f32, confidence: f64, dist_coeffs: Option<&[f64]>,
                         method: Option<SolvePnPMethod>) -> Result<Option<PNPRANSACSolution>, MatError> {
    let _ = dist_coeffs;
    let n = point_correspondences.len();
    let mut obj = Vec::with_capacity(n * 3);
    let mut img = Vec::with_capacity(n * 2);
    for c in point_correspondences {
        obj.extend_from_slice(&[c.obj_point.x, c.obj_point.y, c.obj_point.z]);
        img.extend_from_slice(&[c.img_point.x, c.img_point.y]);
    }
    let mut k = [0f64; 9];
    for r in 0..3 {
        for c in 0..3 {
            k[(r * 3 + c) as usize] = *camera_intrinsic.at_2d(r, c)?;
        }
    }
    let (mut rvec, mut tvec) = ([0f64; 3], [0f64; 3]);
    let mut inliers = vec![0i32; n.max(1)];
    let (mut n_inliers, mut found) = (0i32, 0i32);
    // SAFETY: obj / img hold n x 3 / n x 2 f64; the inlier buffer holds n entries as promised by inliers_cap
    status(unsafe {
        sys::dunk_pnp_ransac(sys::ctx(), obj.as_ptr(), img.as_ptr(), n as i32, k.as_ptr(), iter_count, reproj_thres, confidence,
                             method.unwrap_or(SolvePnPMethod::SOLVEPNP_EPNP) as i32, rvec.as_mut_ptr(), tvec.as_mut_ptr(), inliers.as_mut_ptr(),
                             n as i32, &mut n_inliers, &mut found)
    })?;
    if found == 0 {
        return Ok(None);
    }
    let col = |v: &[f64; 3]| Cmat::<f64>::from_2d_slice(&[[v[0]], [v[1]], [v[2]]]);
    let rows: Vec<[i32; 1]> = inliers[..n_inliers as usize].iter().map(|&i| [i]).collect();
    Ok(Some(PNPRANSACSolution { rvec: col(&rvec)?, tvec: col(&tvec)?, inliers: Cmat::<i32>::from_2d_slice(&rows)? }))
}
