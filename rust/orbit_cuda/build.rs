fn main() {
    // ORBIT_B200_LIB_DIR = <repo>/orbit_b200/lib
    if let Ok(dir) = std::env::var("ORBIT_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=orbit_b200");
}
