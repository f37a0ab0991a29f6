//! `ImageDatabase` (feature_database/src/imagedb.rs:14-77) over the shard's host-resident `ref_image` table.
use crate::{check, models, DbConn, DbError};
use dunk_b200_sys as sys;

/// imagedb.rs:8-11
pub enum Image<'a> {
    One(models::InsertImage<'a>),
    Multiple(Vec<models::InsertImage<'a>>),
}

fn insert(conn: &mut DbConn, im: &models::InsertImage) -> Result<i32, DbError> {
    let mut id = 0i32;
    // SAFETY: live handle, valid out pointer
    check(unsafe { sys::dunk_db_create_image(conn.db, *im.x_start, *im.y_start, *im.x_end, *im.y_end, *im.level_of_detail, &mut id) })?;
    Ok(id)
}

fn ids(conn: &mut DbConn, use_box: i32, b: [i32; 4], lod: i32) -> Result<Vec<i32>, DbError> {
    // SAFETY: live handle
    let cap = unsafe { sys::dunk_db_image_count(conn.db) }.max(1);
    let mut out = vec![0i32; cap as usize];
    let mut n = 0i32;
    // SAFETY: `out` holds `cap` ids
    check(unsafe { sys::dunk_db_find_images(conn.db, use_box, b[0], b[1], b[2], b[3], lod, out.as_mut_ptr(), cap, &mut n) })?;
    out.truncate(n as usize);
    Ok(out)
}

impl ImageDatabase for Image<'_> {
    /// imagedb.rs:14-30 — returns the id of the (first) inserted row; ids are 1-based like a Postgres SERIAL
    fn create_image(conn: &mut DbConn, input_image: Image) -> Result<i32, DbError> {
        match input_image {
            Image::One(im) => insert(conn, &im),
            Image::Multiple(v) => {
                let all: Result<Vec<i32>, DbError> = v.iter().map(|im| insert(conn, im)).collect();
                all?.first().copied().ok_or(DbError::NotFound)
            }
        }
    }

    /// imagedb.rs:32-37
    fn read_image_from_id(conn: &mut DbConn, id: i32) -> Result<models::Image, DbError> {
        let mut r = sys::DunkImage { id: 0, x_start: 0, y_start: 0, x_end: 0, y_end: 0, level_of_detail: 0 };
        // SAFETY: live handle, valid out pointer
        check(unsafe { sys::dunk_db_read_image(conn.db, id, &mut r) })?;
        Ok(models::Image { id: r.id, x_start: r.x_start, y_start: r.y_start, x_end: r.x_end, y_end: r.y_end, level_of_detail: r.level_of_detail })
    }

    /// imagedb.rs:39-56 — images of the LoD whose extent intersects the box
    fn find_images_from_dimensions(conn: &mut DbConn, x_start: i32, y_start: i32, x_end: i32, y_end: i32,
                                   level_of_detail: i32) -> Result<Vec<i32>, DbError> {
        ids(conn, 1, [x_start, y_start, x_end, y_end], level_of_detail)
    }

    /// imagedb.rs:58-66
    fn find_images_from_lod(conn: &mut DbConn, level_of_detail: i32) -> Result<Vec<i32>, DbError> {
        ids(conn, 0, [0; 4], level_of_detail)
    }

    /// imagedb.rs:68-73 — the HBM store is append-only (a shard is rebuilt, not edited in place)
    fn delete_image(_conn: &mut DbConn, _id: i32) -> Result<(), DbError> {
        Err(DbError::Store(sys::DUNK_ERR_BAD_ARG, "the HBM store is append-only: clear and rebuild the shard".into()))
    }
}

/// imagedb.rs:93-110
pub trait ImageDatabase {
    fn create_image(conn: &mut DbConn, image: Image) -> Result<i32, DbError>;
    fn read_image_from_id(conn: &mut DbConn, id: i32) -> Result<models::Image, DbError>;
    fn find_images_from_dimensions(conn: &mut DbConn, x_start: i32, y_start: i32, x_end: i32, y_end: i32,
                                   level_of_detail: i32) -> Result<Vec<i32>, DbError>;
    fn find_images_from_lod(conn: &mut DbConn, level_of_detail: i32) -> Result<Vec<i32>, DbError>;
    fn delete_image(conn: &mut DbConn, id: i32) -> Result<(), DbError>;
}
