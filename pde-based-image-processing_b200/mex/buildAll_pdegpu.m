function buildAll_pdegpu(repo, outdir)
%BUILDALL_PDEGPU  Build the 13 hot-path MEX files of PDE-based-image-processing against libpdegpu.
%
%   buildAll_pdegpu                      % repo = two levels above this file, outdir = './build'
%   buildAll_pdegpu(repo, outdir)
%
% Replaces lines 5-26 of the reference's mex/buildAll.m (one `mex` command per gateway, each compiling
% mex/source/<Name>.c with a library translation unit): the gateways in
% <repo>/pde-based-image-processing_b200/gateways/ have the same names and Matlab-visible signatures, and the
% arithmetic runs in libpdegpu.so (CUDA, built by `python pde-based-image-processing_b200/build.py`).
% The level-set / RANSAC MEX files (buildAll.m:28-40) are not on the path libpdegpu covers: build them with the
% reference's own script. Afterwards runme.m and every matlab/*/*.m driver run unchanged:
%
%   addpath(outdir); runme
%
% Works with Matlab (`mex`) and GNU Octave (`mkoctfile --mex`). Neither exists in the image libpdegpu is developed
% in, so this file is untested there; what is tested are the same gateway sources compiled against
% gateways/mex_shim/mex.h and driven the way the interpreter drives them (tests/test_abi.py, tests/test_gpu_*.py).

here = fileparts(mfilename('fullpath'));
if nargin < 1 || isempty(repo),   repo = fullfile(here, '..', '..'); end
if nargin < 2 || isempty(outdir), outdir = fullfile(pwd, 'build'); end
pkg = fullfile(repo, 'pde-based-image-processing_b200');
gw  = fullfile(pkg, 'gateways');
lib = fullfile(pkg, 'libpdegpu.so');
if ~exist(lib, 'file')
    error('pdegpu:nolib', '%s not found: run `python %s` first (needs nvcc)', lib, fullfile(pkg, 'build.py'));
end
if ~exist(outdir, 'dir'), mkdir(outdir); end

names = {'Oflow_lhs_elin4_2d', 'Oflow_lhs_llin4_2d', 'Oflow_sor_elin4_2d', 'Oflow_sor_llin4_2d', 'Oflow_sor_llin8_2d', ...
         'Disp_sor_llin4_2d', 'Disp_sor_llin_sym4_2d', 'DdiffWeights', 'FstDerivatives5', 'SndDerivatives5', ...
         'BilinInterp_2d', 'PDEsolver4', 'PDEsolver8'};
isOctave = exist('OCTAVE_VERSION', 'builtin') ~= 0;
for k = 1:numel(names)
    src = fullfile(gw, [names{k} '.c']);
    ctx = fullfile(gw, 'gw_ctx.c');
    if isOctave
        cmd = sprintf('mkoctfile --mex -I"%s" -I"%s" "%s" "%s" -L"%s" -lpdegpu -Wl,-rpath,"%s" -o "%s"', ...
                      fullfile(repo, 'include'), gw, src, ctx, pkg, pkg, fullfile(outdir, [names{k} '.mex']));
        [status, msg] = system(cmd);
        if status ~= 0, error('pdegpu:build', '%s\n%s', cmd, msg); end
    else
        mex('-R2018a', ['-I' fullfile(repo, 'include')], ['-I' gw], src, ctx, ['-L' pkg], '-lpdegpu', ...
            ['LDFLAGS=$LDFLAGS -Wl,-rpath,' pkg], '-outdir', outdir);
    end
    fprintf('built %s\n', names{k});
end
fprintf('libpdegpu gateways in %s. GPU: PDEGPU_DEVICE (default 0).\n', outdir);
end
