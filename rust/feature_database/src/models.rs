//! Row types of the reference's `ref_image` and `keypoint` tables (feature_database/src/models.rs:5-55), without the
//! diesel derives: the same field names and types, so callers that build `InsertKeypoint` / read `Keypoint` compile
//! unchanged.

/// models.rs:5-15
#[derive(Clone, Copy, Debug)]
pub struct Image {
    pub id: i32,
    pub x_start: i32,
    pub y_start: i32,
    pub x_end: i32,
    pub y_end: i32,
    pub level_of_detail: i32,
}

/// models.rs:17-25
#[derive(Clone, Copy, Debug)]
pub struct InsertImage<'a> {
    pub x_start: &'a i32,
    pub y_start: &'a i32,
    pub x_end: &'a i32,
    pub y_end: &'a i32,
    pub level_of_detail: &'a i32,
}

/// models.rs:27-41
#[derive(Clone, Debug)]
pub struct Keypoint {
    pub id: i32,
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

/// models.rs:43-55
#[derive(Clone, Debug)]
pub struct InsertKeypoint<'a> {
    pub x_coord: &'a f32,
    pub y_coord: &'a f32,
    pub size: &'a f32,
    pub angle: &'a f32,
    pub response: &'a f32,
    pub octave: &'a i32,
    pub class_id: &'a i32,
    pub descriptor: &'a [u8],
    pub image_id: &'a i32,
}
