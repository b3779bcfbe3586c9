/* Implementation of oracle/ref_shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * calib3d calls and Mat::inv are forwarded to the real OpenCV living inside the `cv2` Python module (4.13); the GIL is
 * taken around each forwarded call, so the reference's OpenMP threads may call in concurrently (cv2 releases the GIL
 * while it computes).  The module that binds the reference's entry points (ref_module.cpp) releases the GIL first.
 *
 * Optional fast path (off by default, `esac_ref.set_native_project(True)`): projectPoints WITHOUT Jacobian runs as a
 * native loop that restates OpenCV's per-point arithmetic (fp64 transform, z ? 1/z : 1, fx*x + cx, round to float) with
 * the rotation matrix still obtained from cv2.Rodrigues.  tests/test_ref_pin.py checks it is bit-identical to
 * cv2.projectPoints; it exists because the Python binding of projectPoints always computes the 2N x 15 Jacobian, which the
 * C++ reference does not pay for in esac_forward -- used only for the CPU timing baseline.
 */
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>

#include <atomic>

#include "opencv2/opencv.hpp"

namespace py = pybind11;

namespace esac_ref_shim {
std::atomic<int> native_project{0};
std::atomic<long> n_solvepnp{0}, n_solvepnp_fail{0}, n_project{0}, n_rodrigues{0}, n_inv{0}, n_cv_error{0};
}  // namespace esac_ref_shim

namespace {

py::object& cv2_module() {
    static py::object* m = nullptr;  /* leaked on purpose: destroyed interpreters must not run ~object */
    if (!m) m = new py::object(py::module_::import("cv2"));
    return *m;
}

py::array_t<double> to_np_f64(const cv::Mat& m) {
    py::array_t<double> a({m.rows, m.cols});
    auto w = a.mutable_unchecked<2>();
    for (int r = 0; r < m.rows; r++) for (int c = 0; c < m.cols; c++) w(r, c) = m.get(r, c);
    return a;
}

py::array_t<float> to_np_f32(const cv::Mat& m) {
    py::array_t<float> a({m.rows, m.cols});
    auto w = a.mutable_unchecked<2>();
    for (int r = 0; r < m.rows; r++) for (int c = 0; c < m.cols; c++) w(r, c) = (float)m.get(r, c);
    return a;
}

/* camera matrix keeps its own depth (the reference passes CV_32F) */
py::object cam_to_np(const cv::Mat& K) {
    if (K.depth() == CV_32F) return to_np_f32(K);
    return to_np_f64(K);
}

cv::Mat from_np_f64(const py::array& arr) {
    py::array_t<double, py::array::c_style | py::array::forcecast> a(arr);
    int rows, cols;
    if (a.ndim() == 1) { rows = (int)a.shape(0); cols = 1; }
    else if (a.ndim() == 2) { rows = (int)a.shape(0); cols = (int)a.shape(1); }
    else throw std::runtime_error("shim: unexpected array rank from cv2");
    cv::Mat m(rows, cols, CV_64F);
    const double* p = a.data();
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<double>(r, c) = p[(size_t)r * cols + c];
    return m;
}

py::array_t<float> points3_to_np(const std::vector<cv::Point3f>& v) {
    py::array_t<float> a({(py::ssize_t)v.size(), (py::ssize_t)1, (py::ssize_t)3});
    std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(cv::Point3f));
    return a;
}

py::array_t<float> points2_to_np(const std::vector<cv::Point2f>& v) {
    py::array_t<float> a({(py::ssize_t)v.size(), (py::ssize_t)1, (py::ssize_t)2});
    std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(cv::Point2f));
    return a;
}

void np_to_points2(const py::array& arr, std::vector<cv::Point2f>& out) {
    py::array_t<float, py::array::c_style> a(arr);
    if (arr.dtype().kind() != 'f' || arr.dtype().itemsize() != 4) throw std::runtime_error("shim: projectPoints did not return float32");
    size_t n = (size_t)a.size() / 2;
    out.resize(n);
    std::memcpy(out.data(), a.data(), n * sizeof(cv::Point2f));
}

void check_binary(const cv::Mat& a, const cv::Mat& b, const char* what) {
    if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error(std::string("shim cv::Mat ") + what + ": size mismatch");
}

int result_type(const cv::Mat& a, const cv::Mat& b) { return (a.depth() == CV_32F && b.depth() == CV_32F) ? CV_32F : CV_64F; }

}  // namespace

namespace cv {

/* ------------------------------------------------------------------ core arithmetic (plain IEEE loops) */
Mat Mat::t() const {
    Mat o(cols, rows, type_);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) std::memcpy(o.ptr(c) + (size_t)r * esz_, ptr(r) + (size_t)c * esz_, esz_);
    return o;
}

Mat& Mat::operator+=(const Mat& b) {
    check_binary(*this, b, "+=");
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, get(r, c) + b.get(r, c));
    return *this;
}
Mat& Mat::operator-=(const Mat& b) {
    check_binary(*this, b, "-=");
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, get(r, c) - b.get(r, c));
    return *this;
}
Mat& Mat::operator*=(double s) {
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, get(r, c) * s);
    return *this;
}

Mat operator*(const Mat& a, const Mat& b) {
    if (a.cols != b.rows) throw std::runtime_error("shim cv::Mat *: inner dimensions differ");
    if (a.depth() != CV_64F || b.depth() != CV_64F) throw std::runtime_error("shim cv::Mat *: only CV_64F products are used by the reference");
    Mat o(a.rows, b.cols, CV_64F);
    const int K = a.cols;
    for (int r = 0; r < a.rows; r++) {
        for (int c = 0; c < b.cols; c++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += a.at<double>(r, k) * b.at<double>(k, c);
            o.at<double>(r, c) = s;
        }
    }
    return o;
}
Mat operator+(const Mat& a, const Mat& b) {
    check_binary(a, b, "+");
    Mat o(a.rows, a.cols, result_type(a, b));
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.set(r, c, a.get(r, c) + b.get(r, c));
    return o;
}
Mat operator-(const Mat& a, const Mat& b) {
    check_binary(a, b, "-");
    Mat o(a.rows, a.cols, result_type(a, b));
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.set(r, c, a.get(r, c) - b.get(r, c));
    return o;
}
Mat operator-(const Mat& a) {
    Mat o(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.set(r, c, -a.get(r, c));
    return o;
}
Mat operator*(const Mat& a, double s) {
    Mat o(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.set(r, c, a.get(r, c) * s);
    return o;
}
Mat operator*(double s, const Mat& a) { return a * s; }
Mat operator/(const Mat& a, double s) {
    Mat o(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.set(r, c, a.get(r, c) / s);
    return o;
}
Mat operator!=(const Mat& a, const Mat& b) {
    check_binary(a, b, "!=");
    Mat o(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.at<uchar>(r, c) = (a.get(r, c) != b.get(r, c)) ? 255 : 0;
    return o;
}
Scalar trace(const Mat& m) {
    double s = 0;
    for (int i = 0; i < std::min(m.rows, m.cols); i++) s += m.get(i, i);
    return Scalar(s);
}
Scalar sum(const Mat& m) {
    double s = 0;
    for (int r = 0; r < m.rows; r++) for (int c = 0; c < m.cols; c++) s += m.get(r, c);
    return Scalar(s);
}
double norm(const Mat& m) {
    double s = 0;
    for (int r = 0; r < m.rows; r++) for (int c = 0; c < m.cols; c++) { double v = m.get(r, c); s += v * v; }
    return std::sqrt(s);
}

/* ------------------------------------------------------------------ real OpenCV through cv2 */
Mat Mat::inv(int method) const {
    esac_ref_shim::n_inv++;
    py::gil_scoped_acquire gil;
    py::tuple res = cv2_module().attr("invert")(to_np_f64(*this), py::none(), method);
    Mat out = from_np_f64(res[1].cast<py::array>());
    if (depth() == CV_32F) { Mat f; out.convertTo(f, CV_32F); return f; }
    return out;
}

bool solvePnP(const std::vector<Point3f>& objectPoints, const std::vector<Point2f>& imagePoints, const Mat& cameraMatrix,
              const Mat& distCoeffs, Mat& rvec, Mat& tvec, bool useExtrinsicGuess, int flags) {
    (void)distCoeffs; /* the reference always passes cv::Mat() */
    esac_ref_shim::n_solvepnp++;
    py::gil_scoped_acquire gil;
    try {
        py::object r_in = rvec.empty() ? py::object(py::none()) : py::object(to_np_f64(rvec));
        py::object t_in = tvec.empty() ? py::object(py::none()) : py::object(to_np_f64(tvec));
        py::tuple res = cv2_module().attr("solvePnP")(points3_to_np(objectPoints), points2_to_np(imagePoints), cam_to_np(cameraMatrix),
                                                      py::none(), r_in, t_in, useExtrinsicGuess, flags);
        bool ok = res[0].cast<bool>();
        if (ok) {
            rvec = from_np_f64(res[1].cast<py::array>());
            tvec = from_np_f64(res[2].cast<py::array>());
        } else {
            esac_ref_shim::n_solvepnp_fail++;
        }
        return ok;
    } catch (py::error_already_set& e) {
        /* a cv::Exception inside the C++ reference would terminate the OpenMP region; report it as a failed solve */
        esac_ref_shim::n_cv_error++;
        esac_ref_shim::n_solvepnp_fail++;
        e.restore();
        PyErr_Clear();
        return false;
    }
}

static void rotation_of(const Mat& rvec, double R[9]) {
    py::tuple res = cv2_module().attr("Rodrigues")(to_np_f64(rvec));
    Mat Rm = from_np_f64(res[0].cast<py::array>());
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = Rm.at<double>(r, c);
}

void projectPoints(const std::vector<Point3f>& objectPoints, const Mat& rvec, const Mat& tvec, const Mat& cameraMatrix,
                   const Mat& distCoeffs, std::vector<Point2f>& imagePoints) {
    (void)distCoeffs;
    esac_ref_shim::n_project++;
    if (esac_ref_shim::native_project.load()) {
        double R[9], t[3];
        {
            py::gil_scoped_acquire gil;
            rotation_of(rvec, R);
        }
        for (int i = 0; i < 3; i++) t[i] = tvec.get(tvec.rows == 1 ? 0 : i, tvec.rows == 1 ? i : 0);
        const double fx = cameraMatrix.get(0, 0), fy = cameraMatrix.get(1, 1), cx = cameraMatrix.get(0, 2), cy = cameraMatrix.get(1, 2);
        const size_t n = objectPoints.size();
        imagePoints.resize(n);
        for (size_t i = 0; i < n; i++) {
            const double X = objectPoints[i].x, Y = objectPoints[i].y, Z = objectPoints[i].z;
            double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
            double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
            double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
            z = z ? 1. / z : 1;
            x *= z;
            y *= z;
            imagePoints[i].x = (float)(x * fx + cx);
            imagePoints[i].y = (float)(y * fy + cy);
        }
        return;
    }
    py::gil_scoped_acquire gil;
    py::tuple res = cv2_module().attr("projectPoints")(points3_to_np(objectPoints), to_np_f64(rvec), to_np_f64(tvec),
                                                       cam_to_np(cameraMatrix), py::none());
    np_to_points2(res[0].cast<py::array>(), imagePoints);
}

void projectPoints(const std::vector<Point3f>& objectPoints, const Mat& rvec, const Mat& tvec, const Mat& cameraMatrix,
                   const Mat& distCoeffs, std::vector<Point2f>& imagePoints, Mat& jacobian) {
    (void)distCoeffs;
    esac_ref_shim::n_project++;
    py::gil_scoped_acquire gil;
    py::tuple res = cv2_module().attr("projectPoints")(points3_to_np(objectPoints), to_np_f64(rvec), to_np_f64(tvec),
                                                       cam_to_np(cameraMatrix), py::none());
    np_to_points2(res[0].cast<py::array>(), imagePoints);
    py::array_t<double, py::array::c_style | py::array::forcecast> J(res[1].cast<py::array>());
    Mat out((int)J.shape(0), (int)J.shape(1), CV_64F);
    std::memcpy(out.ptr(0), J.data(), (size_t)J.size() * sizeof(double));
    jacobian = out;
}

void Rodrigues(const Mat& src, Mat& dst) {
    esac_ref_shim::n_rodrigues++;
    py::gil_scoped_acquire gil;
    py::tuple res = cv2_module().attr("Rodrigues")(to_np_f64(src));
    dst = from_np_f64(res[0].cast<py::array>());
}

void Rodrigues(const Mat& src, Mat& dst, Mat& jacobian) {
    esac_ref_shim::n_rodrigues++;
    py::gil_scoped_acquire gil;
    py::tuple res = cv2_module().attr("Rodrigues")(to_np_f64(src));
    dst = from_np_f64(res[0].cast<py::array>());
    jacobian = from_np_f64(res[1].cast<py::array>());
}

}  // namespace cv
