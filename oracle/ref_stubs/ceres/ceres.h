// TEST INFRASTRUCTURE ONLY (oracle/): the slice of the Ceres 1.12 public interface that the reference's factor sources derive
// from (ceres::CostFunction / SizedCostFunction / LossFunction / LocalParameterization), so that
// /root/reference/vins_estimator/src/factor/* compiles unmodified into oracle/_ref/. No solver lives here.
// CauchyLoss follows Ceres' published definition rho(s) = b log(1 + s/b), b = a^2 (loss_function.cc).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace ceres {

typedef int int32;

class CostFunction {
  public:
    CostFunction() : num_residuals_(0) {}
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    const std::vector<int32> &parameter_block_sizes() const { return parameter_block_sizes_; }
    int num_residuals() const { return num_residuals_; }

  protected:
    std::vector<int32> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }

  private:
    std::vector<int32> parameter_block_sizes_;
    int num_residuals_;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
  public:
    SizedCostFunction() {
        set_num_residuals(kNumResiduals);
        const int sizes[] = {Ns...};
        for (int s : sizes)
            if (s > 0) mutable_parameter_block_sizes()->push_back(s);
    }
    virtual ~SizedCostFunction() {}
};

class LossFunction {
  public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

class CauchyLoss : public LossFunction {
  public:
    explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
    virtual void Evaluate(double s, double rho[3]) const {
        const double sum = 1.0 + s * c_;
        const double inv = 1.0 / sum;
        rho[0] = b_ * std::log(sum);
        rho[1] = std::max(std::numeric_limits<double>::min(), inv);
        rho[2] = -c_ * (inv * inv);
    }

  private:
    const double b_, c_;
};

class HuberLoss : public LossFunction {
  public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    virtual void Evaluate(double s, double rho[3]) const {
        if (s > b_) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a_ * r - b_;
            rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else {
            rho[0] = s;
            rho[1] = 1.0;
            rho[2] = 0.0;
        }
    }

  private:
    const double a_, b_;
};

class LocalParameterization {
  public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};

}  // namespace ceres
