//! `geotransform` / `elevation` (feature_database/src/elevationdb.rs:12-244): the two GDAL geotransforms and the
//! elevation raster live in HBM (`dunk_elevation`); `get_world_coordinates` — pixel -> geotransform -> nearest
//! elevation sample -> EPSG:4326 -> EPSG:4978 (ECEF) — is `dunk_world_coordinates`, batched on the device.
//! GDAL dataset reading (`add_elevation_data(conn, &Dataset)`) is replaced by a raster slice.
use crate::{check, DbConn, DbError};
use dunk_b200_sys as sys;

/// elevationdb.rs:6-10 (the GDAL variant has no counterpart: nothing here calls GDAL)
#[derive(Debug)]
pub enum Errors {
    Diesel(DbError),
}

/// a GDAL geotransform: [x0, dx, rx, y0, ry, dy]
pub type GeoTransform = [f64; 6];

fn handle(conn: &mut DbConn) -> Result<*mut sys::DunkElevation, DbError> {
    if conn.elevation.is_null() {
        let gd = conn.gt_dataset.ok_or(DbError::NotFound)?; // read_geotransform(conn, "dataset") fails -> NotFound
        // the reference falls back to height 0 when the "elevation" transform is absent (elevationdb.rs:76-79)
        let with_dem = conn.gt_elevation.is_some() && conn.heights.is_some();
        let mut e = std::ptr::null_mut();
        let rc = match (&conn.gt_elevation, &conn.heights) {
            (Some(ge), Some((h, xs, ys))) if with_dem =>
                // SAFETY: 6-element transforms, xs * ys heights
                unsafe { sys::dunk_elevation_create(sys::ctx(), gd.as_ptr(), ge.as_ptr(), h.as_ptr(), *xs, *ys, &mut e) },
            _ =>
                // SAFETY: NULL elevation transform / raster = height 0
                unsafe { sys::dunk_elevation_create(sys::ctx(), gd.as_ptr(), std::ptr::null(), std::ptr::null(), 0, 0, &mut e) },
        };
        check(rc)?;
        conn.elevation = e;
    }
    Ok(conn.elevation)
}

fn invalidate(conn: &mut DbConn) {
    if !conn.elevation.is_null() {
        // SAFETY: created by dunk_elevation_create, destroyed once
        unsafe { sys::dunk_elevation_destroy(conn.elevation) };
        conn.elevation = std::ptr::null_mut();
    }
}

pub mod geotransform {
    use super::*;

    /// elevationdb.rs:21-36 — `name` is "dataset" or "elevation"
    pub fn create_geotransform(conn: &mut DbConn, name: &str, transform: GeoTransform) -> Result<(), DbError> {
        match name {
            "dataset" => conn.gt_dataset = Some(transform),
            "elevation" => conn.gt_elevation = Some(transform),
            _ => return Err(DbError::Store(sys::DUNK_ERR_BAD_ARG, format!("unknown geotransform name {name:?}"))),
        }
        invalidate(conn);
        Ok(())
    }

    /// elevationdb.rs:64-90 — reference-image pixel (x, y) -> ECEF metres.  A missing elevation sample is the
    /// reference's diesel NotFound.
    pub fn get_world_coordinates(conn: &mut DbConn, x: f64, y: f64) -> Result<(f64, f64, f64), Errors> {
        let v = get_world_coordinates_batch(conn, &[x], &[y])?;
        Ok(v[0])
    }

    /// the batched form the pipeline uses: one launch for all matched reference keypoints of a frame
    pub fn get_world_coordinates_batch(conn: &mut DbConn, x: &[f64], y: &[f64]) -> Result<Vec<(f64, f64, f64)>, Errors> {
        assert_eq!(x.len(), y.len());
        let e = handle(conn).map_err(Errors::Diesel)?;
        let mut xyz = vec![0f64; x.len() * 3];
        let mut missing = 0i32;
        // SAFETY: x / y hold n values, xyz holds 3 n
        check(unsafe { sys::dunk_world_coordinates(e, x.as_ptr(), y.as_ptr(), x.len() as i64, xyz.as_mut_ptr(), &mut missing) })
            .map_err(Errors::Diesel)?;
        if missing > 0 {
            return Err(Errors::Diesel(DbError::NotFound));
        }
        Ok(xyz.chunks_exact(3).map(|c| (c[0], c[1], c[2])).collect())
    }
}

pub mod elevation {
    use super::*;

    /// elevationdb.rs:190-232 with the GDAL read factored out: `heights` is band 1 of the elevation dataset in row-major
    /// order (`y_size` rows of `x_size` samples) — the `elevation` table in row-id order plus `elevation_properties`
    pub fn add_elevation_data(conn: &mut DbConn, heights: &[f64], x_size: i32, y_size: i32) -> Result<(), Errors> {
        if heights.len() != (x_size as usize) * (y_size as usize) {
            return Err(Errors::Diesel(DbError::Store(sys::DUNK_ERR_VEC_LENGTH, "heights.len() != x_size * y_size".into())));
        }
        conn.heights = Some((heights.to_vec(), x_size, y_size));
        invalidate(conn);
        Ok(())
    }

    /// elevationdb.rs:234-246 — row id = round(y) * x_size + round(x) + 1 (f64::round: half away from zero)
    pub fn get_elevation(conn: &mut DbConn, x: f64, y: f64) -> Result<f64, DbError> {
        let (h, xs, ys) = conn.heights.as_ref().ok_or(DbError::NotFound)?;
        let idx = y.round() as i64 * *xs as i64 + x.round() as i64;
        if idx < 0 || idx >= *xs as i64 * *ys as i64 {
            return Err(DbError::NotFound);
        }
        Ok(h[idx as usize])
    }
}
