function varargout=dbat_mex(varargin) %#ok<STOUT,INUSD>
%DBAT_MEX Gateway to libdbatgpu (B200 bundle-adjustment inner loop).
%
%   See mex/dbat_mex.c for the commands.  This stub is shadowed by the
%   compiled MEX file, following the convention of icpc_mex.m.

error('Mex file not found.');
