//! Drop-in for the READ / LOAD side of the reference crate `feature_database`: the
//! `keypoint` and `ref_image` tables live in HBM as SoA columns of a `dunk_db` shard instead of Postgres, the
//! geotransform + elevation tables as a `dunk_elevation` handle.  Trait and function names, argument order and result
//! types follow the reference; the connection argument is `&mut DbConn` (the HBM store) where the reference takes
//! `&mut diesel::pg::PgConnection`, and `DbError` stands where it returns `diesel::result::Error` — a caller swaps
//! those two `use` lines.  Postgres / diesel I/O, migrations and GDAL raster reading are out of scope.
pub mod elevationdb;
pub mod imagedb;
pub mod keypointdb;
pub mod models;

use dunk_b200_sys as sys;

/// what the reference surfaces as `diesel::result::Error`
#[derive(Debug)]
pub enum DbError {
    /// diesel::result::Error::NotFound
    NotFound,
    /// any other failure of the store: (status code of include/dunk_b200.h, message)
    Store(i32, String),
}

pub(crate) fn check(rc: i32) -> Result<(), DbError> {
    match rc {
        0 => Ok(()),
        sys::DUNK_ERR_OUT_OF_RANGE => Err(DbError::NotFound),
        _ => Err(DbError::Store(rc, sys::last_error())),
    }
}

/// The HBM-resident store a "connection" addresses: one `dunk_db` shard (keypoint + ref_image tables) and, once
/// `create_geotransform` / `add_elevation_data` have been called, one `dunk_elevation` handle.
pub struct DbConn {
    pub(crate) db: *mut sys::DunkDb,
    pub(crate) gt_dataset: Option<[f64; 6]>,
    pub(crate) gt_elevation: Option<[f64; 6]>,
    pub(crate) heights: Option<(Vec<f64>, i32, i32)>,
    pub(crate) elevation: *mut sys::DunkElevation,
}

// SAFETY: the library serialises mutation of a dunk_db under its own mutex; the handles are plain pointers
unsafe impl Send for DbConn {}

impl DbConn {
    /// stands where the reference calls `PgConnection::establish(url)`: an empty store of `capacity_rows` rows
    pub fn establish(capacity_rows: i64) -> Result<Self, DbError> {
        let mut db = std::ptr::null_mut();
        // SAFETY: valid out pointer
        check(unsafe { sys::dunk_db_create(sys::ctx(), capacity_rows, sys::DUNK_DESC_BYTES, &mut db) })?;
        Ok(DbConn { db, gt_dataset: None, gt_elevation: None, heights: None, elevation: std::ptr::null_mut() })
    }

    /// re-load a flat dump written by `save` (the 3 GB reference DB is built once, then loaded at H2D speed)
    pub fn load(path: &str, min_capacity_rows: i64) -> Result<Self, DbError> {
        let c = std::ffi::CString::new(path).map_err(|_| DbError::Store(sys::DUNK_ERR_BAD_ARG, "path contains NUL".into()))?;
        let mut db = std::ptr::null_mut();
        // SAFETY: NUL-terminated path, valid out pointer
        check(unsafe { sys::dunk_db_load(sys::ctx(), c.as_ptr(), min_capacity_rows, &mut db) })?;
        Ok(DbConn { db, gt_dataset: None, gt_elevation: None, heights: None, elevation: std::ptr::null_mut() })
    }

    pub fn save(&mut self, path: &str) -> Result<(), DbError> {
        let c = std::ffi::CString::new(path).map_err(|_| DbError::Store(sys::DUNK_ERR_BAD_ARG, "path contains NUL".into()))?;
        // SAFETY: live handle, NUL-terminated path
        check(unsafe { sys::dunk_db_save(self.db, c.as_ptr()) })
    }

    /// the raw shard handle, for `dunk_db_match` / `dunk_register_frames*`
    pub fn raw(&self) -> *mut sys::DunkDb {
        self.db
    }

    pub fn len(&self) -> i64 {
        // SAFETY: live handle
        unsafe { sys::dunk_db_size(self.db) }
    }

    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
}

impl Drop for DbConn {
    fn drop(&mut self) {
        // SAFETY: the handles were created by the library and are destroyed exactly once
        unsafe {
            if !self.elevation.is_null() {
                sys::dunk_elevation_destroy(self.elevation);
            }
            sys::dunk_db_destroy(self.db);
        }
    }
}
