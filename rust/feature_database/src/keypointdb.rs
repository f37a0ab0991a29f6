//! `KeypointDatabase` (feature_database/src/keypointdb.rs:14-109) over the HBM shard: inserts append SoA rows, keyed
//! reads run on the device (predicate -> stable descending sort on `response` -> gather, `dunk_db_select`).
use crate::{check, models, DbConn, DbError};
use dunk_b200_sys as sys;

/// keypointdb.rs:7-10
pub enum Keypoint<'a> {
    One(models::InsertKeypoint<'a>),
    Multiple(Vec<models::InsertKeypoint<'a>>),
}

/// keypointdb.rs:12
const OPENCV_KEYPOINT_LIMIT: i64 = 2_i64.pow(18) - 1;
const DESC: usize = sys::DUNK_DESC_BYTES as usize;

fn append(conn: &mut DbConn, rows: &[models::InsertKeypoint]) -> Result<(), DbError> {
    let mut desc = Vec::with_capacity(rows.len() * DESC);
    let mut kps = Vec::with_capacity(rows.len());
    let mut ids = Vec::with_capacity(rows.len());
    for r in rows {
        if r.descriptor.len() != DESC {
            return Err(DbError::Store(sys::DUNK_ERR_ASSERT, format!("descriptor of {} bytes, {} expected", r.descriptor.len(), DESC)));
        }
        desc.extend_from_slice(r.descriptor);
        kps.push(sys::DunkKeyPoint { x: *r.x_coord, y: *r.y_coord, size: *r.size, angle: *r.angle, response: *r.response,
                                     octave: *r.octave, class_id: *r.class_id });
        ids.push(*r.image_id);
    }
    // SAFETY: the three columns hold rows.len() entries each
    check(unsafe { sys::dunk_db_append(conn.db, desc.as_ptr(), kps.as_ptr(), ids.as_ptr(), rows.len() as i64) })
}

/// rows [0, n) of a shard as models::Keypoint
fn read_all(db: *mut sys::DunkDb) -> Result<Vec<models::Keypoint>, DbError> {
    // SAFETY: live handle
    let n = unsafe { sys::dunk_db_size(db) } as usize;
    let mut desc = vec![0u8; n * DESC];
    let mut kps = vec![sys::DunkKeyPoint { x: 0.0, y: 0.0, size: 0.0, angle: 0.0, response: 0.0, octave: 0, class_id: 0 }; n];
    let (mut img, mut ids) = (vec![0i32; n], vec![0i32; n]);
    if n > 0 {
        // SAFETY: every output column holds n entries
        check(unsafe { sys::dunk_db_read(db, 0, n as i64, desc.as_mut_ptr(), kps.as_mut_ptr(), img.as_mut_ptr()) })?;
        check(unsafe { sys::dunk_db_read_ids(db, 0, n as i64, ids.as_mut_ptr()) })?;
    }
    Ok((0..n)
        .map(|i| models::Keypoint { id: ids[i], x_coord: kps[i].x, y_coord: kps[i].y, size: kps[i].size, angle: kps[i].angle,
                                     response: kps[i].response, octave: kps[i].octave, class_id: kps[i].class_id,
                                     descriptor: desc[i * DESC..(i + 1) * DESC].to_vec(), image_id: img[i] })
        .collect())
}

fn select(conn: &mut DbConn, f: sys::DunkRowFilter) -> Result<Vec<models::Keypoint>, DbError> {
    let mut sub = std::ptr::null_mut();
    // SAFETY: live handle, filter by reference, valid out pointer
    check(unsafe { sys::dunk_db_select(conn.db, &f, OPENCV_KEYPOINT_LIMIT, &mut sub) })?;
    let rows = read_all(sub);
    // SAFETY: `sub` was created by the call above
    unsafe { sys::dunk_db_destroy(sub) };
    rows
}

impl<'a> KeypointDatabase for Keypoint<'a> {
    /// keypointdb.rs:15-26
    fn create_keypoint(conn: &mut DbConn, input_keypoint: Keypoint) -> Result<(), DbError> {
        match input_keypoint {
            Keypoint::One(k) => append(conn, &[k]),
            Keypoint::Multiple(v) => append(conn, &v),
        }
    }

    /// keypointdb.rs:28-36 — the `id` column is 1 + row index
    fn read_keypoint_from_id(conn: &mut DbConn, id: i32) -> Result<models::Keypoint, DbError> {
        if id < 1 || id as i64 > conn.len() {
            return Err(DbError::NotFound);
        }
        let mut desc = vec![0u8; DESC];
        let mut kp = sys::DunkKeyPoint { x: 0.0, y: 0.0, size: 0.0, angle: 0.0, response: 0.0, octave: 0, class_id: 0 };
        let mut img = 0i32;
        // SAFETY: one-row outputs
        check(unsafe { sys::dunk_db_read(conn.db, (id - 1) as i64, 1, desc.as_mut_ptr(), &mut kp, &mut img) })?;
        Ok(models::Keypoint { id, x_coord: kp.x, y_coord: kp.y, size: kp.size, angle: kp.angle, response: kp.response, octave: kp.octave,
                              class_id: kp.class_id, descriptor: desc, image_id: img })
    }

    /// keypointdb.rs:38-49 — WHERE image_id = ? ORDER BY response DESC LIMIT 2^18 - 1
    fn read_keypoints_from_image_id(conn: &mut DbConn, image_id: i32) -> Result<Vec<models::Keypoint>, DbError> {
        select(conn, sys::DunkRowFilter { image_id, level_of_detail: -1, use_box: 0, x_start: 0.0, y_start: 0.0, x_end: 0.0, y_end: 0.0 })
    }

    /// keypointdb.rs:51-66 — JOIN ref_image WHERE level_of_detail = ?
    fn read_keypoints_from_lod(conn: &mut DbConn, level_of_detail: i32) -> Result<Vec<models::Keypoint>, DbError> {
        select(conn, sys::DunkRowFilter { image_id: -1, level_of_detail, use_box: 0, x_start: 0.0, y_start: 0.0, x_end: 0.0, y_end: 0.0 })
    }

    /// keypointdb.rs:68-90 — the box bounds are floor()ed / ceil()ed and inclusive (done by the select kernel)
    fn read_keypoints_from_coordinates(conn: &mut DbConn, x_start: f32, y_start: f32, x_end: f32, y_end: f32,
                                       level_of_detail: i32) -> Result<Vec<models::Keypoint>, DbError> {
        select(conn, sys::DunkRowFilter { image_id: -1, level_of_detail, use_box: 1, x_start, y_start, x_end, y_end })
    }

    /// keypointdb.rs:92-97 — the HBM store is append-only
    fn delete_keypoint(_conn: &mut DbConn, _id: i32) -> Result<(), DbError> {
        Err(DbError::Store(sys::DUNK_ERR_BAD_ARG, "the HBM store is append-only: clear and rebuild the shard".into()))
    }
}

/// keypointdb.rs:111-140
pub trait KeypointDatabase {
    fn create_keypoint(conn: &mut DbConn, input_keypoint: Keypoint) -> Result<(), DbError>;
    fn read_keypoint_from_id(conn: &mut DbConn, id: i32) -> Result<models::Keypoint, DbError>;
    fn read_keypoints_from_image_id(conn: &mut DbConn, image_id: i32) -> Result<Vec<models::Keypoint>, DbError>;
    fn read_keypoints_from_lod(conn: &mut DbConn, level_of_detail: i32) -> Result<Vec<models::Keypoint>, DbError>;
    fn read_keypoints_from_coordinates(conn: &mut DbConn, x_start: f32, y_start: f32, x_end: f32, y_end: f32,
                                       level_of_detail: i32) -> Result<Vec<models::Keypoint>, DbError>;
    fn delete_keypoint(conn: &mut DbConn, id: i32) -> Result<(), DbError>;
}
